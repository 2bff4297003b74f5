// Host-only timing of TreeHost::build_from_parents (what polee_set_tree does before its upload).
//   nvcc -O3 -std=c++17 -I polee_b200/csrc -I include tools/ubench/tree_host_time.cu polee_b200/csrc/build/tree_host.o \
//        polee_b200/csrc/build/mem_cache.o -lcudart -o /tmp/tree_host_time && /tmp/tree_host_time tree.bin
// tree.bin: int32 node_parent_idxs[2n-1] followed by node_js[2n-1]
#include <chrono>
#include <cstdio>
#include <vector>
#include "common.cuh"
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<int32_t> buf(bytes / 4);
    if (fread(buf.data(), 4, buf.size(), f) != buf.size()) return 1;
    const int64_t N = (int64_t)buf.size() / 2, n = (N + 1) / 2;
    for (int rep = 0; rep < 6; ++rep) {
        polee::TreeHost th;
        auto t0 = std::chrono::steady_clock::now();
        std::string e = th.build_from_parents(n, buf.data(), buf.data() + N, 512);
        auto t1 = std::chrono::steady_clock::now();
        printf("rep %d: %.2f ms  err='%s' preorder=%d dfs_bwd=%d\n", rep, std::chrono::duration<double, std::milli>(t1 - t0).count(),
               e.c_str(), (int)th.preorder, (int)th.dfs_bwd);
    }
    return 0;
}
