"""The Rust side of the boundary (rust/annembed_cuda_sys.rs, source only: no Rust toolchain in the image) must declare
exactly what include/annembed_cuda.h declares: every function with the same number and kind of arguments, every struct
with the same fields in the same order and of the same width, every status code and flag with the same value.
Both files are parsed as text; the test fails on drift."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "annembed_cuda.h")
RUST = os.path.join(ROOT, "rust", "annembed_cuda_sys.rs")

C2R = {"uint32_t": "u32", "uint64_t": "u64", "uint8_t": "u8", "double": "f64", "float": "f32", "int": "c_int", "char": "c_char"}


def _strip_c(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def _strip_rust(src):
    return re.sub(r"//[^\n]*", "", src)


def c_arg_type(arg):
    """'const uint64_t *row_ptr' -> '*const u64'; 'uint8_t handles[128]' -> '*mut u8'; 'int device' -> 'c_int'."""
    arg = arg.strip()
    const = "const " in arg
    array = "[" in arg
    stars = arg.count("*") + (1 if array else 0)
    base = re.sub(r"\[.*?\]", "", arg.replace("const", "").replace("*", " ")).split()
    name_dropped = base[:-1] if len(base) > 1 else base            # last token is the parameter name
    t = " ".join(name_dropped)
    t = C2R.get(t, t)
    for _ in range(stars):
        t = ("*const " if const else "*mut ") + t
        const = False if stars > 1 else const                       # 'T **' = *mut *mut T
    return t.replace("*mut *const", "*mut *mut")


def c_functions():
    src = _strip_c(open(HEADER).read())
    out = {}
    for ret, name, args in re.findall(r"\b(int|const char \*)\s*(annembed_cuda_[a-z_0-9]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        out[name] = [c_arg_type(a) for a in args.split(",") if a.strip() and a.strip() != "void"]
    return out


def rust_functions():
    src = _strip_rust(open(RUST).read())
    out = {}
    for name, args in re.findall(r"pub fn (annembed_cuda_[a-z_0-9]+)\s*\(([^;]*?)\)\s*->", src, flags=re.S):
        out[name] = [a.split(":", 1)[1].strip() for a in args.split(",") if ":" in a]
    return out


def c_structs():
    src = _strip_c(open(HEADER).read())
    out = {}
    for body, name in re.findall(r"typedef struct [a-z_]+ \{(.*?)\}\s*([a-z_]+)\s*;", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(\w+)\s+(\w+)(\[(\d+)\])?$", decl)
            assert m, decl
            t = C2R[m.group(1)]
            fields.append((m.group(2), f"[{t}; {m.group(4)}]" if m.group(4) else t))
        out[name] = fields
    return out


def rust_structs():
    src = _strip_rust(open(RUST).read())
    out = {}
    for name, body in re.findall(r"pub struct (annembed_cuda_[a-z_]+)\s*\{(.*?)\}", src, flags=re.S):
        out[name] = [(f.split(":")[0].replace("pub", "").strip(), f.split(":")[1].strip()) for f in body.split(",") if ":" in f]
    return out


def test_every_function_is_declared_with_the_same_signature():
    c, r = c_functions(), rust_functions()
    assert len(c) >= 30
    assert set(c) == set(r), f"functions differ: {set(c) ^ set(r)}"
    for name in c:
        cargs = [a.replace("annembed_cuda_ctx", "annembed_cuda_ctx") for a in c[name]]
        assert cargs == r[name], (name, cargs, r[name])


def test_structs_match_field_by_field():
    c, r = c_structs(), rust_structs()
    for name in ("annembed_cuda_params", "annembed_cuda_stats", "annembed_cuda_quality"):
        assert c[name] == r[name], (name, c[name], r[name])


def test_status_codes_and_flags_match():
    csrc = _strip_c(open(HEADER).read())
    rsrc = _strip_rust(open(RUST).read())
    c_consts = {k: int(v) for k, v in re.findall(r"\b(ANNEMBED_ERR_[A-Z_]+|ANNEMBED_OK)\s*=\s*(\d+)", csrc)}
    c_consts.update({k: int(v) for k, v in re.findall(r"#define (ANNEMBED_FLAG_[A-Z0-9_]+|ANNEMBED_CUDA_ABI_VERSION) (\d+)u?", csrc)})
    r_consts = {k: int(v) for k, v in re.findall(r"pub const (ANNEMBED_[A-Z0-9_]+): \w+ = (\d+);", rsrc)}
    assert len(c_consts) >= 14
    assert c_consts == r_consts, set(c_consts.items()) ^ set(r_consts.items())


def test_csr_file_format_matches_the_python_reader():
    """rust/kgraph_csr.rs and annembed_b200/kgraph.py describe the same file: magic, version, field order."""
    from annembed_b200 import kgraph
    rs = open(os.path.join(ROOT, "rust", "kgraph_csr.rs")).read()
    assert kgraph.MAGIC == re.search(r'MAGIC: &\[u8; 8\] = b"(\w+)"', rs).group(1).encode()
    assert "VERSION: u32 = 1" in rs
    order = [rs.index(f"for v in &self.{f}") for f in ("row_ptr", "col", "dist", "data_id")]
    assert order == sorted(order)
