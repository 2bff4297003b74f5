// hclust.cu -- the tree heuristic of PolyaTreeTransform(X, :cluster) on the host.
//
// src/hclust.jl:193-319 (hclust, hclust_join_edges!) + :361-389 (order_nodes): transcripts are ordered by the median
// index of their compatible reads, every transcript is compared with its K = 25 successors (Jaccard similarity of the
// sorted read sets), subtrees that share the most reads are joined greedily (the union's read set replaces the two),
// what remains without any shared read is joined smallest first, and the tree is serialised in DFS order, right
// branch first -- the (node_parent_idxs, node_js) pair polee_set_tree takes and .prep.h5 stores.
//
// This is host code (pointer chasing over growing sets; it runs once per sample, before the device work) and it is
// PARITY-UNPINNED by construction: the reference breaks similarity ties by the internal order of
// DataStructures.jl's binary heap and by the iteration order of a Julia Dict, neither of which is specified, and its
// own test tree predates the current heuristic.  The policy here is explicit instead: among equal Float32
// similarities the edge with the smaller (j1, j2) wins; equal sizes in the remainder queue: smaller node id first;
// a node's neighbour list keeps first occurrences only.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <vector>

#include "../../include/polee_b200.h"

namespace {

using Set = std::vector<uint32_t>;

size_t intersection_size(const Set &a, const Set &b) {  // hclust.jl:111-131
    if (a.empty() || b.empty() || a.front() > b.back() || a.back() < b.front()) return 0;
    size_t i = 0, j = 0, c = 0;
    while (i < a.size() && j < b.size()) {
        if (a[i] < b[j])
            ++i;
        else if (a[i] > b[j])
            ++j;
        else {
            ++i; ++j; ++c;
        }
    }
    return c;
}

float jaccard(const Set &a, const Set &b) {  // read_set_relative_intersection_size, :139-149; stored as Float32 (:84)
    if (a.empty() && b.empty()) return 0.0f;
    const size_t c = intersection_size(a, b);
    return (float)((double)c / (double)(a.size() + b.size() - c));
}

struct Edge {
    uint32_t j1, j2;
    float sim;
};
struct EdgeLess {  // max-heap: highest similarity on top; ties: smaller (j1, j2) on top
    bool operator()(const Edge &a, const Edge &b) const {
        if (a.sim != b.sim) return a.sim < b.sim;
        if (a.j1 != b.j1) return a.j1 > b.j1;
        return a.j2 > b.j2;
    }
};
struct Sized {
    uint32_t j;
    uint64_t size;
};
struct SizedGreater {  // min-heap by size; ties: smaller id on top
    bool operator()(const Sized &a, const Sized &b) const { return a.size != b.size ? a.size > b.size : a.j > b.j; }
};

}  // namespace

extern "C" int polee_hclust(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval, int32_t *node_parent_idxs,
                            int32_t *node_js) {
    if (!colptr || !rowval || !node_parent_idxs || !node_js || n < 1 || m < 1 || colptr[0] != 1) return POLEE_EINVAL;
    for (int64_t j = 0; j < n; ++j)
        if (colptr[j + 1] < colptr[j]) return POLEE_EINVAL;
    const int64_t K = 25;                       // :199
    const int64_t N = 2 * n - 1;
    // transcripts by median compatible read index (:203-211); sortperm is stable
    std::vector<uint32_t> med(n);
    for (int64_t j = 0; j < n; ++j)
        med[j] = colptr[j] == colptr[j + 1] ? 0u : rowval[((uint64_t)colptr[j] + colptr[j + 1]) / 2 - 1];
    std::vector<uint32_t> idxs(n);
    std::iota(idxs.begin(), idxs.end(), 0u);
    std::stable_sort(idxs.begin(), idxs.end(), [&](uint32_t a, uint32_t b) { return med[a] < med[b]; });

    // nodes 0..n-1 leaves (in that order), n.. internal
    std::vector<Set> sets(N);
    std::vector<int32_t> left(N, -1), right(N, -1), leaf_tx(N, 0);
    std::vector<char> dead(N, 0), live(N, 0);
    std::vector<std::vector<uint32_t>> nbr(N);
    for (int64_t j = 0; j < n; ++j) {
        const uint32_t t = idxs[j];
        sets[j].assign(rowval + (colptr[t] - 1), rowval + (colptr[t + 1] - 1));
        leaf_tx[j] = (int32_t)t + 1;
        live[j] = 1;
    }
    std::priority_queue<Edge, std::vector<Edge>, EdgeLess> queue;
    auto add_nbr = [&](uint32_t a, uint32_t b) {
        auto &v = nbr[a];
        if (std::find(v.begin(), v.end(), b) == v.end()) v.push_back(b);
    };
    for (int64_t j1 = 0; j1 < n; ++j1)  // :224-233
        for (int64_t j2 = j1 + 1; j2 <= std::min(j1 + K, n - 1); ++j2) {
            const float s = jaccard(sets[j1], sets[j2]);
            if (s > 0) queue.push(Edge{(uint32_t)j1, (uint32_t)j2, s});
            add_nbr((uint32_t)j1, (uint32_t)j2);
            add_nbr((uint32_t)j2, (uint32_t)j1);
        }
    uint32_t next = (uint32_t)n;
    while (!queue.empty()) {  // hclust_join_edges!, :266-319
        const Edge e = queue.top();
        queue.pop();
        if (dead[e.j1] || dead[e.j2]) continue;
        const uint32_t k = next++;
        Set merged;
        merged.reserve(sets[e.j1].size() + sets[e.j2].size());
        std::set_union(sets[e.j1].begin(), sets[e.j1].end(), sets[e.j2].begin(), sets[e.j2].end(), std::back_inserter(merged));
        sets[k].swap(merged);
        left[k] = (int32_t)e.j1;
        right[k] = (int32_t)e.j2;
        live[k] = 1;
        for (uint32_t j : {e.j1, e.j2}) {
            Set().swap(sets[j]);
            dead[j] = 1;
            live[j] = 0;
        }
        const uint32_t pairs[2][2] = {{e.j1, e.j2}, {e.j2, e.j1}};
        for (const auto &pr : pairs) {
            const std::vector<uint32_t> list = nbr[pr[0]];
            for (uint32_t l : list) {
                if (l == pr[1] || dead[l]) continue;
                const float s = jaccard(sets[l], sets[k]);
                if (s != 0) queue.push(Edge{l, k, s});
                add_nbr(l, k);
                add_nbr(k, l);
            }
            std::vector<uint32_t>().swap(nbr[pr[0]]);
        }
    }
    // what shares no read with anything: smallest first (:243-262)
    std::priority_queue<Sized, std::vector<Sized>, SizedGreater> rest;
    for (uint32_t j = 0; j < next; ++j)
        if (live[j]) rest.push(Sized{j, 1 + (uint64_t)sets[j].size()});
    while (rest.size() > 1) {
        const Sized a = rest.top(); rest.pop();
        const Sized b = rest.top(); rest.pop();
        const uint32_t k = next++;
        left[k] = (int32_t)a.j;
        right[k] = (int32_t)b.j;
        rest.push(Sized{k, a.size + b.size});
    }
    if ((int64_t)next != N) return POLEE_EBADTREE;
    const uint32_t root = rest.top().j;
    // order_nodes (:361-389): DFS, the right child is popped first
    std::vector<uint32_t> stack{root};
    std::vector<int32_t> parent_of(N, 0);
    int64_t pos = 0;
    while (!stack.empty()) {
        const uint32_t v = stack.back();
        stack.pop_back();
        node_parent_idxs[pos] = parent_of[v];
        node_js[pos] = left[v] < 0 ? leaf_tx[v] : 0;
        ++pos;
        if (left[v] >= 0) {
            parent_of[left[v]] = (int32_t)pos;  // 1-based index of v
            parent_of[right[v]] = (int32_t)pos;
            stack.push_back((uint32_t)left[v]);
            stack.push_back((uint32_t)right[v]);
        }
    }
    return pos == N ? POLEE_OK : POLEE_EBADTREE;
}
