"""Output format of the reference's callers (N4): CSV dumps of the embedding with Rust's `{:.5e}`
formatting -- ≙ write_csv_array2 / write_csv_labeled_array2 (/root/reference/src/tools/io.rs:23-67)."""
from __future__ import annotations

import numpy as np


def rust_lower_exp(x: float, precision: int = 5) -> str:
    """Rust's `format!("{:.5e}", x)` for an f32: `1.23457e0`, `-9.87654e-3` (no '+', no zero-padded exponent)."""
    x = float(np.float32(x))
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    mant, exp = f"{x:.{precision}e}".split("e")
    return f"{mant}e{int(exp)}"


def write_csv_array2(path: str, mat: np.ndarray) -> int:
    """≙ write_csv_array2 (tools/io.rs:47-67): one row per line, comma separated."""
    mat = np.asarray(mat)
    with open(path, "w", newline="") as f:
        for row in mat:
            f.write(",".join(rust_lower_exp(v) for v in row) + "\n")
    return 1


def write_csv_labeled_array2(path: str, labels, mat: np.ndarray) -> int:
    """≙ write_csv_labeled_array2 (tools/io.rs:23-44): label first, then the row."""
    mat = np.asarray(mat)
    with open(path, "w", newline="") as f:
        for lab, row in zip(labels, mat):
            f.write(",".join([str(lab)] + [rust_lower_exp(v) for v in row]) + "\n")
    return 1
