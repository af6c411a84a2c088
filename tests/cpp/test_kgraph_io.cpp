// CPU-only driver of the C++ mirror's ANNKGCSR reader / writer (include/annembed_embedder.hpp): reads argv[1], checks
// the CSR invariants, writes the same graph to argv[2]; with "bad" as argv[2] expects read_csr to throw.
// tests/test_host.py compares the bytes with annembed_b200/kgraph.py's.
#include <cstdio>
#include <cstring>
#include "../../include/annembed_embedder.hpp"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    try {
        annembed::KGraph g = annembed::read_csr(argv[1]);
        if (!std::strcmp(argv[2], "bad")) return 1;          // a malformed file must not load
        const size_t n = g.get_nb_nodes();
        for (size_t i = 0; i < n; i++) {
            if (g.row_ptr[i] > g.row_ptr[i + 1]) return 3;
            for (uint64_t m = g.row_ptr[i] + 1; m < g.row_ptr[i + 1]; m++) if (g.dist[m] < g.dist[m - 1]) return 4;   // kgraph.rs:508-509
        }
        std::printf("n=%zu E=%zu max_nbng=%zu first_id=%llu\n", n, g.col.size(), g.get_max_nbng(), (unsigned long long)g.get_data_id_from_idx(0));
        annembed::write_csr(argv[2], g);
    } catch (const std::exception &e) {
        std::printf("error: %s\n", e.what());
        return std::strcmp(argv[2], "bad") ? 5 : 0;
    }
    return 0;
}
