// factorize.cu -- exact row de-duplication of a likelihood matrix on the device.
//
// tools/exact-factorization.jl:31-68 of the reference hashes every row of X -- the pair (transcript ids, Float32
// values) -- into a Dict, keeps one copy per distinct row with its multiplicity `counts`, and feeds the compressed
// matrix to the factored likelihood (likelihood.jl:59-85, polee_set_matrix_csc(..., ks)).  Identical fragments are
// common in deep samples, and every byte the sparse kernels stream is per row, so this is the one algorithmic saving
// that does not change the result.  Here: CSR by a stable radix sort, a 64-bit hash per row, a stable sort by hash,
// an exact comparison of neighbours (a hash collision can only split a group, never merge two), groups numbered by
// first occurrence.  The reference's row order is the iteration order of a Julia Dict (unpinned); first occurrence is
// this implementation's choice, everything else (the set of rows, the counts, column order inside a row) is exact.
#include <algorithm>
#include <cub/cub.cuh>
#include <vector>

#include "common.cuh"

namespace polee {
namespace {

__global__ void k_fz_count(const uint32_t *__restrict__ rowval, int64_t nnz, uint32_t *row_len, int64_t m, int *bad) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r = rowval[e] - 1u;
        if (r >= (uint64_t)m)
            *bad = 1;
        else
            atomicAdd(&row_len[r], 1u);
    }
}

__global__ void k_fz_expand(const uint32_t *__restrict__ colptr, int64_t n, const uint32_t *__restrict__ rowval, int64_t nnz,
                            uint32_t *col_of, uint32_t *key_row, uint32_t *val_e) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)colptr[mid] - 1 <= e)
                lo = mid;
            else
                hi = mid;
        }
        col_of[e] = (uint32_t)lo;
        key_row[e] = rowval[e] - 1u;
        val_e[e] = (uint32_t)e;
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    return h ^ (h >> 33);
}

// one thread per row: hash of (length, transcript ids, value bits) in the row's ascending-transcript order
__global__ void k_fz_hash(int64_t m, const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ a_csc,
                          const uint32_t *__restrict__ col_of, const float *__restrict__ nzval, uint64_t *hash,
                          uint32_t *row_id) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = row_ptr[i], e = row_ptr[i + 1];
        uint64_t h = mix64(0x243f6a8885a308d3ull, e - b);
        for (uint32_t q = b; q < e; ++q) {
            const uint32_t src = a_csc[q];
            h = mix64(h, ((uint64_t)col_of[src] << 32) | (uint64_t)__float_as_uint(nzval[src]));
        }
        hash[i] = h;
        row_id[i] = (uint32_t)i;
    }
}

// p = position in hash order: does the row start a new group (differs from its predecessor)?
__global__ void k_fz_new(int64_t m, const uint64_t *__restrict__ hash_s, const uint32_t *__restrict__ row_s,
                         const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ a_csc,
                         const uint32_t *__restrict__ col_of, const float *__restrict__ nzval, uint32_t *is_new) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m; p += (int64_t)gridDim.x * blockDim.x) {
        bool nw = p == 0 || hash_s[p] != hash_s[p - 1];
        if (!nw) {
            const uint32_t i = row_s[p], j = row_s[p - 1];
            const uint32_t bi = row_ptr[i], bj = row_ptr[j], li = row_ptr[i + 1] - bi;
            nw = li != row_ptr[j + 1] - bj;
            for (uint32_t t = 0; t < li && !nw; ++t) {
                const uint32_t si = a_csc[bi + t], sj = a_csc[bj + t];
                nw = col_of[si] != col_of[sj] || __float_as_uint(nzval[si]) != __float_as_uint(nzval[sj]);
            }
        }
        is_new[p] = nw ? 1u : 0u;
    }
}

// group leader = first row of the group in hash order (the stable sort keeps original order inside a group, so it is
// the group's first occurrence); group sizes
__global__ void k_fz_groups(int64_t m, const uint32_t *__restrict__ is_new, const uint32_t *__restrict__ gid_incl,
                            const uint32_t *__restrict__ row_s, uint32_t *leader, uint32_t *gsize) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t g = gid_incl[p] - 1u;
        if (is_new[p]) leader[g] = row_s[p];
        atomicAdd(&gsize[g], 1u);
    }
}

// groups sorted by leader row = order of first occurrence: unique index of every leader row, counts
__global__ void k_fz_number(uint32_t n_groups, const uint32_t *__restrict__ leader_s, const uint32_t *__restrict__ group_s,
                            const uint32_t *__restrict__ gsize, uint32_t *uidx_of_row, int64_t *counts) {
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_groups; u += gridDim.x * blockDim.x) {
        uidx_of_row[leader_s[u]] = u + 1u;  // 0 = the row is not a leader
        counts[u] = (int64_t)gsize[group_s[u]];
    }
}

__global__ void k_fz_keep(const uint32_t *__restrict__ rowval, int64_t nnz, const uint32_t *__restrict__ uidx_of_row,
                          uint32_t *keep) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e <= nnz; e += (int64_t)gridDim.x * blockDim.x)
        keep[e] = (e < nnz && uidx_of_row[rowval[e] - 1u] != 0u) ? 1u : 0u;
}

__global__ void k_fz_emit(const uint32_t *__restrict__ rowval, const float *__restrict__ nzval, int64_t nnz,
                          const uint32_t *__restrict__ keep, const uint32_t *__restrict__ kept_before,
                          const uint32_t *__restrict__ uidx_of_row, uint32_t *rowval_out, float *nzval_out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
        if (keep[e]) {
            rowval_out[kept_before[e]] = uidx_of_row[rowval[e] - 1u];  // already 1-based
            nzval_out[kept_before[e]] = nzval[e];
        }
}

__global__ void k_fz_colptr(const uint32_t *__restrict__ colptr, int64_t n, const uint32_t *__restrict__ kept_before,
                            uint32_t *colptr_out) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j <= n; j += (int64_t)gridDim.x * blockDim.x)
        colptr_out[j] = kept_before[colptr[j] - 1u] + 1u;
}

struct Scratch {
    std::vector<void *> ptrs;
    ~Scratch() {
        for (void *p : ptrs) polee::dfree(p);
    }
    template <typename T>
    cudaError_t alloc(T **p, size_t count) {
        cudaError_t e = polee::dmalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

int bits_for64(uint64_t maxval) {
    int b = 1;
    while (b < 64 && (maxval >> b) != 0) ++b;
    return b;
}

}  // namespace
}  // namespace polee

using namespace polee;

#define FZ(expr)                                  \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) return POLEE_ECUDA; \
    } while (0)

extern "C" int polee_exact_factorization(int32_t device, int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                                         const float *nzval, int64_t *m_unique, uint32_t *colptr_out, uint32_t *rowval_out,
                                         float *nzval_out, int64_t *counts_out, int64_t *nnz_out) {
    if (!colptr || !rowval || !nzval || !m_unique || !colptr_out || !rowval_out || !nzval_out || !counts_out || !nnz_out)
        return POLEE_EINVAL;
    // the CUB sorts and scans take int item counts
    if (m < 1 || n < 1 || colptr[0] != 1 || m > (int64_t)INT32_MAX || (int64_t)colptr[n] - 1 > (int64_t)INT32_MAX) return POLEE_EINVAL;
    for (int64_t j = 0; j < n; ++j)
        if (colptr[j + 1] < colptr[j]) return POLEE_EINVAL;
    const int64_t nnz = (int64_t)colptr[n] - 1;
    FZ(cudaSetDevice(device));
    cudaStream_t st;
    FZ(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    } guard{st};
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int TPB = 256;
    auto grid_for = [&](int64_t work) { return (int)std::max<int64_t>(1, std::min<int64_t>((work + TPB - 1) / TPB, (int64_t)sms * 32)); };
    Scratch sc;
    uint32_t *d_colptr, *d_rowval, *row_len, *row_ptr, *col_of, *key_row, *val_e, *row_sorted, *a_csc;
    float *d_nzval;
    int *d_bad;
    FZ(sc.alloc(&d_colptr, n + 1)); FZ(sc.alloc(&d_rowval, nnz)); FZ(sc.alloc(&d_nzval, nnz));
    FZ(sc.alloc(&row_len, m + 1)); FZ(sc.alloc(&row_ptr, m + 1)); FZ(sc.alloc(&col_of, nnz)); FZ(sc.alloc(&key_row, nnz));
    FZ(sc.alloc(&val_e, nnz)); FZ(sc.alloc(&row_sorted, nnz)); FZ(sc.alloc(&a_csc, nnz)); FZ(sc.alloc(&d_bad, 1));
    FZ(cudaMemcpyAsync(d_colptr, colptr, sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, st));
    FZ(cudaMemcpyAsync(d_rowval, rowval, sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice, st));
    FZ(cudaMemcpyAsync(d_nzval, nzval, sizeof(float) * nnz, cudaMemcpyHostToDevice, st));
    FZ(cudaMemsetAsync(row_len, 0, sizeof(uint32_t) * (m + 1), st));
    FZ(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    size_t need = 0, tmp_bytes = 0;
    void *d_tmp = nullptr;
    auto ensure_tmp = [&](size_t bytes) -> cudaError_t {
        if (bytes <= tmp_bytes) return cudaSuccess;
        tmp_bytes = bytes + bytes / 8;
        return sc.alloc((char **)&d_tmp, tmp_bytes);
    };
    // ---- CSR order (row, ascending transcript)
    if (nnz > 0) {
        k_fz_count<<<grid_for(nnz), TPB, 0, st>>>(d_rowval, nnz, row_len, m, d_bad);
        k_fz_expand<<<grid_for(nnz), TPB, 0, st>>>(d_colptr, n, d_rowval, nnz, col_of, key_row, val_e);
        FZ(cub::DeviceRadixSort::SortPairs(nullptr, need, key_row, row_sorted, val_e, a_csc, (int)nnz, 0, 32, st));
        FZ(ensure_tmp(need));
        FZ(cub::DeviceRadixSort::SortPairs(d_tmp, need, key_row, row_sorted, val_e, a_csc, (int)nnz, 0, bits_for64((uint64_t)m), st));
    }
    FZ(cub::DeviceScan::ExclusiveSum(nullptr, need, row_len, row_ptr, (int)(m + 1), st));
    FZ(ensure_tmp(need));
    FZ(cub::DeviceScan::ExclusiveSum(d_tmp, need, row_len, row_ptr, (int)(m + 1), st));
    int bad = 0;
    FZ(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    // ---- hash, sort by hash (stable), neighbours compared exactly
    uint64_t *hash, *hash_s;
    uint32_t *row_id, *row_s, *is_new, *gid_incl;
    FZ(sc.alloc(&hash, m)); FZ(sc.alloc(&hash_s, m)); FZ(sc.alloc(&row_id, m)); FZ(sc.alloc(&row_s, m));
    FZ(sc.alloc(&is_new, m)); FZ(sc.alloc(&gid_incl, m));
    k_fz_hash<<<grid_for(m), TPB, 0, st>>>(m, row_ptr, a_csc, col_of, d_nzval, hash, row_id);
    FZ(cub::DeviceRadixSort::SortPairs(nullptr, need, hash, hash_s, row_id, row_s, (int)m, 0, 64, st));
    FZ(ensure_tmp(need));
    FZ(cub::DeviceRadixSort::SortPairs(d_tmp, need, hash, hash_s, row_id, row_s, (int)m, 0, 64, st));
    k_fz_new<<<grid_for(m), TPB, 0, st>>>(m, hash_s, row_s, row_ptr, a_csc, col_of, d_nzval, is_new);
    FZ(cub::DeviceScan::InclusiveSum(nullptr, need, is_new, gid_incl, (int)m, st));
    FZ(ensure_tmp(need));
    FZ(cub::DeviceScan::InclusiveSum(d_tmp, need, is_new, gid_incl, (int)m, st));
    uint32_t n_groups = 0;
    FZ(cudaMemcpyAsync(&n_groups, gid_incl + (m - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    FZ(cudaStreamSynchronize(st));
    if (bad) return POLEE_EINVAL;
    // ---- groups in order of first occurrence
    uint32_t *leader, *gsize, *group_id, *leader_s, *group_s, *uidx_of_row, *keep, *kept_before;
    int64_t *d_counts;
    FZ(sc.alloc(&leader, n_groups)); FZ(sc.alloc(&gsize, n_groups)); FZ(sc.alloc(&group_id, n_groups));
    FZ(sc.alloc(&leader_s, n_groups)); FZ(sc.alloc(&group_s, n_groups)); FZ(sc.alloc(&uidx_of_row, m));
    FZ(sc.alloc(&keep, nnz + 1)); FZ(sc.alloc(&kept_before, nnz + 1)); FZ(sc.alloc(&d_counts, n_groups));
    FZ(cudaMemsetAsync(gsize, 0, sizeof(uint32_t) * n_groups, st));
    FZ(cudaMemsetAsync(uidx_of_row, 0, sizeof(uint32_t) * m, st));
    k_fz_groups<<<grid_for(m), TPB, 0, st>>>(m, is_new, gid_incl, row_s, leader, gsize);
    {
        // iota over groups (reuse the hash kernel's id output pattern with a tiny lambda-free loop)
        std::vector<uint32_t> ids(n_groups);
        for (uint32_t g = 0; g < n_groups; ++g) ids[g] = g;
        FZ(cudaMemcpyAsync(group_id, ids.data(), sizeof(uint32_t) * n_groups, cudaMemcpyHostToDevice, st));
        FZ(cub::DeviceRadixSort::SortPairs(nullptr, need, leader, leader_s, group_id, group_s, (int)n_groups, 0, 32, st));
        FZ(ensure_tmp(need));
        FZ(cub::DeviceRadixSort::SortPairs(d_tmp, need, leader, leader_s, group_id, group_s, (int)n_groups, 0,
                                           bits_for64((uint64_t)m), st));
        FZ(cudaStreamSynchronize(st));  // ids must outlive the copy
    }
    k_fz_number<<<grid_for(n_groups), TPB, 0, st>>>(n_groups, leader_s, group_s, gsize, uidx_of_row, d_counts);
    // ---- compressed CSC: the leaders' entries, rows renumbered (order inside a column is preserved)
    k_fz_keep<<<grid_for(nnz + 1), TPB, 0, st>>>(d_rowval, nnz, uidx_of_row, keep);
    FZ(cub::DeviceScan::ExclusiveSum(nullptr, need, keep, kept_before, (int)(nnz + 1), st));
    FZ(ensure_tmp(need));
    FZ(cub::DeviceScan::ExclusiveSum(d_tmp, need, keep, kept_before, (int)(nnz + 1), st));
    uint32_t nnz_u = 0;
    FZ(cudaMemcpyAsync(&nnz_u, kept_before + nnz, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    uint32_t *d_rowval_out, *d_colptr_out;
    float *d_nzval_out;
    FZ(sc.alloc(&d_rowval_out, nnz)); FZ(sc.alloc(&d_nzval_out, nnz)); FZ(sc.alloc(&d_colptr_out, n + 1));
    if (nnz > 0) k_fz_emit<<<grid_for(nnz), TPB, 0, st>>>(d_rowval, d_nzval, nnz, keep, kept_before, uidx_of_row, d_rowval_out, d_nzval_out);
    k_fz_colptr<<<grid_for(n + 1), TPB, 0, st>>>(d_colptr, n, kept_before, d_colptr_out);
    FZ(cudaStreamSynchronize(st));
    FZ(cudaMemcpyAsync(colptr_out, d_colptr_out, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost, st));
    FZ(cudaMemcpyAsync(rowval_out, d_rowval_out, sizeof(uint32_t) * nnz_u, cudaMemcpyDeviceToHost, st));
    FZ(cudaMemcpyAsync(nzval_out, d_nzval_out, sizeof(float) * nnz_u, cudaMemcpyDeviceToHost, st));
    FZ(cudaMemcpyAsync(counts_out, d_counts, sizeof(int64_t) * n_groups, cudaMemcpyDeviceToHost, st));
    FZ(cudaStreamSynchronize(st));
    FZ(cudaGetLastError());
    *m_unique = (int64_t)n_groups;
    *nnz_out = (int64_t)nnz_u;
    return POLEE_OK;
}
