"""EmbedderParams -- same fields, defaults and meaning as /root/reference/src/embedparams.rs:76-132."""
from __future__ import annotations

from dataclasses import dataclass

# ANNEMBED_FLAG_* of include/annembed_cuda.h
FLAG_GENERIC_EPOCH_KERNEL = 1
FLAG_NO_L2_PERSIST = 2
FLAG_NO_RELABEL = 4
FLAG_REPLAY_IN_EDGES = 8
FLAG_LEGACY_EPOCH_KERNELS = 16
FLAG_BULK_SYNCHRONOUS = 32
FLAG_NODE_ALIAS = 64
FLAG_CP_ASYNC_PIPELINE = 128
FLAG_SECTOR_NEGATIVES = 256


@dataclass
class EmbedderParams:
    asked_dim: int = 2               # embedparams.rs:78,109
    dmap_init: bool = True           # :80,110   (the initial layout is an input of the device path)
    beta: float = 1.0                # :82,111
    b: float = 1.0                   # :85,112
    scale_rho: float = 1.0           # :88,113
    grad_step: float = 2.0           # :90,114
    nb_sampling_by_edge: int = 10    # :92,115
    nb_grad_batch: int = 20          # :94,116
    grad_factor: int = 4             # :98,117
    hierarchy_layer: int = 0         # :100,118
    hubness_weighting: bool = False  # :102,119
    # device-side additions (the reference's RNG is unseeded; see include/annembed_cuda.h)
    mini_epochs_per_batch: int = 0   # 0 -> graded schedule (finest: ceil(nb_sampling_by_edge / 0.15))
    seed: int = 0x5EED
    flags: int = 0
    cell_substeps: int = 0           # cell-resident epoch kernel: mini-epochs per launch (0 -> from the cross-cell edge fraction)

    # setters/getters named as in embedparams.rs:151-183
    def set_dmap_init(self, val: bool): self.dmap_init = val
    def set_nb_gradient_batch(self, nb_batch: int): self.nb_grad_batch = nb_batch
    def set_dim(self, dim: int): self.asked_dim = dim
    def set_nb_edge_sampling(self, nb_sample_by_edge: int): self.nb_sampling_by_edge = nb_sample_by_edge
    def get_dimension(self) -> int: return self.asked_dim
    def set_hierarchy_layer(self, layer: int): self.hierarchy_layer = layer
    def get_hierarchy_layer(self) -> int: return self.hierarchy_layer
