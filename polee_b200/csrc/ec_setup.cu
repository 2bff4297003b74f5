// ec_setup.cu -- builds the equivalence-class layout (common.cuh, "ec") from RNASeqSample.X as Julia stores it
// (SparseMatrixCSC{Float32,UInt32}, 1-based; src/rnaseq_sample.jl:11,499).  Replaces the host-side
// `Xt = SparseMatrixCSC(transpose(X))` of the reference (src/likelihood-approximation.jl:407) for the rows it takes.
//
// All passes run on the device (CUB sorts / scans + small kernels, set-up only):
//   1. CSR order of the entries (stable radix sort by row: ascending transcript inside a row, the reference's
//      accumulation order, src/sparse.jl:14-18);
//   2. a 64-bit hash of every row's transcript-id list, a stable sort of the rows by hash, an exact comparison of
//      neighbours (a collision can only split a class, never merge two) -> classes; classes ordered by first row;
//   3. classes with >= EC_MIN_ROWS rows and 1..EC_MAX_L transcripts are cut into blocks of 32 rows and tasks of
//      <= ec_nbt(L) blocks; one blob per task (header, column ids, partial slots, values lane-major);
//   4. the (task, column) partials are ordered by column (stable radix sort) -> `dest` slots + the second-stage list;
//   5. every other row ("rest") is emitted as a compact CSC of its own for the general layouts (matrix_setup.cu).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>

#include "common.cuh"

namespace polee {

EcRest::~EcRest() {
    polee::dfree(colptr); polee::dfree(rowval); polee::dfree(nzval); polee::dfree(ks);
}

namespace {

__global__ void k_ec_count(const uint32_t *__restrict__ rowval, int64_t nnz, uint32_t *row_len, int64_t m, int *bad) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r = rowval[e] - 1u;
        if (r >= (uint64_t)m)
            *bad = 1;
        else
            atomicAdd(&row_len[r], 1u);
    }
}

__global__ void k_ec_expand(const uint32_t *__restrict__ colptr, int64_t n, const uint32_t *__restrict__ rowval, int64_t nnz,
                            uint32_t *col_of, uint32_t *key_row, uint32_t *val_e) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n;  // largest j with colptr[j] - 1 <= e
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

__device__ __forceinline__ uint64_t ec_mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    return h ^ (h >> 33);
}

// one thread per row: column of every CSR position + hash of (length, transcript ids)
__global__ void k_ec_hash(int64_t m, const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ a_csc,
                          const uint32_t *__restrict__ col_of, uint32_t *col_csr, uint64_t *hash, uint32_t *row_id) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = row_ptr[i], e = row_ptr[i + 1];
        uint64_t h = ec_mix(0x243f6a8885a308d3ull, e - b);
        for (uint32_t q = b; q < e; ++q) {
            const uint32_t c = col_of[a_csc[q]];
            col_csr[q] = c;
            h = ec_mix(h, c);
        }
        hash[i] = h;
        row_id[i] = (uint32_t)i;
    }
}

// p = position in hash order: does the row start a new class?
__global__ void k_ec_new(int64_t m, const uint64_t *__restrict__ hash_s, const uint32_t *__restrict__ row_s,
                         const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col_csr, uint32_t *is_new,
                         uint32_t *rpos) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t i = row_s[p];
        rpos[i] = (uint32_t)p;
        bool nw = p == 0 || hash_s[p] != hash_s[p - 1];
        if (!nw) {
            const uint32_t j = row_s[p - 1];
            const uint32_t bi = row_ptr[i], bj = row_ptr[j], li = row_ptr[i + 1] - bi;
            nw = li != row_ptr[j + 1] - bj;
            for (uint32_t t = 0; t < li && !nw; ++t) nw = col_csr[bi + t] != col_csr[bj + t];
        }
        is_new[p] = nw ? 1u : 0u;
    }
}

__global__ void k_ec_class_start(int64_t m, const uint32_t *__restrict__ is_new, const uint32_t *__restrict__ gid_incl,
                                 const uint32_t *__restrict__ row_s, uint32_t n_classes, uint32_t *cstart, uint32_t *leader,
                                 uint32_t *gid_iota) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m; p += (int64_t)gridDim.x * blockDim.x) {
        if (is_new[p]) {
            const uint32_t g = gid_incl[p] - 1u;
            cstart[g] = (uint32_t)p;
            leader[g] = row_s[p];  // stable sort: the class's first row in the original order
            gid_iota[g] = g;
        }
        if (p == 0) cstart[n_classes] = (uint32_t)m;
    }
}

// c = class in order of first occurrence
__global__ void k_ec_class_meta(uint32_t n_classes, const uint32_t *__restrict__ cls_s, const uint32_t *__restrict__ leader_s,
                                const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ row_len, uint32_t min_rows,
                                uint32_t *order_of_class, uint32_t *cls_L, uint32_t *cls_nt, uint32_t *cls_nb,
                                uint64_t *cls_bytes, uint32_t *cls_parts, uint32_t *cls_rows, uint64_t *cls_ent) {
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c <= n_classes; c += gridDim.x * blockDim.x) {
        if (c == n_classes) {  // sentinels for the exclusive scans
            cls_nt[c] = 0; cls_nb[c] = 0; cls_bytes[c] = 0; cls_parts[c] = 0; cls_rows[c] = 0; cls_ent[c] = 0;
            continue;
        }
        const uint32_t g = cls_s[c];
        order_of_class[g] = c;
        const uint32_t R = cstart[g + 1] - cstart[g], L = row_len[leader_s[c]];
        const bool dense = R >= min_rows && L >= 1u && L <= EC_MAX_L;
        cls_L[c] = dense ? L : 0u;  // 0 marks a class that stays in the general layouts
        uint32_t nb = 0, nt = 0;
        uint64_t bytes = 0;
        if (dense) {
            const uint32_t rpb = 32u / ec_q(L);  // rows per block
            nb = (R + rpb - 1u) / rpb;
            const uint32_t nbt = ec_nbt(L);
            nt = (nb + nbt - 1u) / nbt;
            bytes = (uint64_t)nt * ec_hdr_bytes(L) + (uint64_t)nb * ec_lh(L) * 128ull;
        }
        cls_nt[c] = nt; cls_nb[c] = nb; cls_bytes[c] = bytes; cls_parts[c] = nt * L;
        cls_rows[c] = dense ? R : 0u;
        cls_ent[c] = dense ? (uint64_t)R * L : 0ull;
    }
}

// one thread per task: descriptor, header, column ids, and the column key of each of its partials
__global__ void k_ec_task_init(uint32_t n_tasks, uint32_t n_classes, const uint32_t *__restrict__ task0,
                               const uint32_t *__restrict__ blk0, const uint64_t *__restrict__ byte0,
                               const uint32_t *__restrict__ part0, const uint32_t *__restrict__ cls_L,
                               const uint32_t *__restrict__ cls_nb, const uint32_t *__restrict__ cls_rows,
                               const uint32_t *__restrict__ leader_s, const uint32_t *__restrict__ row_ptr,
                               const uint32_t *__restrict__ col_csr, EcTaskDesc *desc, unsigned char *blob, uint32_t *part_col,
                               uint32_t *part_task) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tasks; t += gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = n_classes;  // largest c with task0[c] <= t (classes without tasks repeat the offset)
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (task0[mid] <= t)
                lo = mid;
            else
                hi = mid;
        }
        const uint32_t c = lo, j = t - task0[c], L = cls_L[c], nbt = ec_nbt(L), nb_cls = cls_nb[c];
        const uint32_t nb = min(nbt, nb_cls - j * nbt);
        const uint64_t off = byte0[c] + (uint64_t)j * ec_task_bytes(L, nbt);
        desc[t] = EcTaskDesc{off, ec_task_bytes(L, nb), 0u};
        EcHdr hd;
        hd.pk = ec_pack(L); hd.nb = nb;
        const uint32_t rpb = 32u / ec_q(L);
        hd.rows = min(nb * rpb, cls_rows[c] - j * nbt * rpb);
        hd.slot0 = (blk0[c] + j * nbt) * 32u;
        unsigned char *b = blob + off;
        *reinterpret_cast<EcHdr *>(b) = hd;
        uint32_t *cols = reinterpret_cast<uint32_t *>(b + 16);
        const uint32_t q0 = row_ptr[leader_s[c]], Lp = ec_q(L) * ec_lh(L);
        for (uint32_t l = 0; l < Lp; ++l) {
            const uint32_t col = l < L ? col_csr[q0 + l] : 0u;  // padding columns point at column 0 (their values are 0)
            cols[l] = col;
            if (l < L) {
                part_col[part0[c] + j * L + l] = col;
                part_task[part0[c] + j * L + l] = t;
            }
        }
    }
}

// the partial array is ordered by column: partial plist[i] (a (task, l) pair) writes slot i
__global__ void k_ec_dest(uint32_t n_parts, const uint32_t *__restrict__ plist, const uint32_t *__restrict__ part_task,
                          const uint32_t *__restrict__ task_part0, const EcTaskDesc *__restrict__ desc, unsigned char *blob) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_parts; i += gridDim.x * blockDim.x) {
        const uint32_t pid = plist[i], t = part_task[pid];
        unsigned char *b = blob + desc[t].off;
        const uint32_t L = ec_pk_l(reinterpret_cast<const EcHdr *>(b)->pk);
        reinterpret_cast<uint32_t *>(b + 16 + 4 * ec_q(L) * ec_lh(L))[pid - task_part0[t]] = i;
    }
}

// processing order of the tasks: by (LH, q) descending, so that the warps of an SM run the same instantiation of the
// kernel's block loop at the same time (its eight instantiations do not fit the instruction cache together) and the
// most expensive tasks come first.  Only the descriptors move; blobs, partial slots and results do not depend on it.
__global__ void k_ec_task_key(uint32_t n_tasks, const EcTaskDesc *__restrict__ desc, const unsigned char *__restrict__ blob,
                              uint32_t *key) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tasks; t += gridDim.x * blockDim.x) {
        const uint32_t pk = reinterpret_cast<const EcHdr *>(blob + desc[t].off)->pk;
        key[t] = 31u - ((ec_pk_lh(pk) - 1u) * 4u + ec_pk_lq(pk));
    }
}

__global__ void k_ec_task_part0(uint32_t n_parts, const uint32_t *__restrict__ part_task, uint32_t *task_part0) {
    for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < n_parts; pid += gridDim.x * blockDim.x)
        if (pid == 0 || part_task[pid] != part_task[pid - 1]) task_part0[part_task[pid]] = pid;
}

// one thread per row (hash order): its row slot, weight, or its "rest" flag
__global__ void k_ec_rows(int64_t m, const uint32_t *__restrict__ row_s, const uint32_t *__restrict__ gid_incl,
                          const uint32_t *__restrict__ order_of_class, const uint32_t *__restrict__ cstart,
                          const uint32_t *__restrict__ cls_L, const uint32_t *__restrict__ blk0,
                          const int64_t *__restrict__ ks, uint32_t *row_of_slot, float *slot_weight, uint32_t *slot_of_row,
                          uint32_t *rest_flag) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= m; p += (int64_t)gridDim.x * blockDim.x) {
        if (p == m) {
            rest_flag[m] = 0;
            continue;
        }
        const uint32_t i = row_s[p], g = gid_incl[p] - 1u, c = order_of_class[g];
        if (cls_L[c] == 0u) {
            rest_flag[i] = 1u;
            slot_of_row[i] = 0xFFFFFFFFu;
        } else {
            rest_flag[i] = 0u;
            const uint32_t rpb = 32u / ec_q(cls_L[c]), rc = (uint32_t)p - cstart[g];  // row of the class
            const uint32_t slot = (blk0[c] + rc / rpb) * 32u + rc % rpb;
            slot_of_row[i] = slot;
            row_of_slot[slot] = i;
            if (slot_weight) slot_weight[slot] = (float)ks[i];
        }
    }
}

// one thread per CSR position: the value goes to its place in the task blob (lane-major, see common.cuh)
__global__ void k_ec_fill(int64_t nnz, const uint32_t *__restrict__ row_sorted, const uint32_t *__restrict__ a_csc,
                          const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ slot_of_row,
                          const uint32_t *__restrict__ rpos, const uint32_t *__restrict__ gid_incl,
                          const uint32_t *__restrict__ order_of_class, const uint32_t *__restrict__ cls_L,
                          const uint32_t *__restrict__ blk0, const uint64_t *__restrict__ byte0,
                          const float *__restrict__ nzval, unsigned char *blob) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t i = row_sorted[q];
        const uint32_t slot = slot_of_row[i];
        if (slot == 0xFFFFFFFFu) continue;
        const uint32_t c = order_of_class[gid_incl[rpos[i]] - 1u];
        const uint32_t L = cls_L[c], qq = ec_q(L), LH = ec_lh(L), nbt = ec_nbt(L);
        const uint32_t blk = slot / 32u - blk0[c], row = slot & 31u;  // row of the block, < 32 / qq
        const uint32_t j = blk / nbt, b = blk % nbt;
        const uint32_t l = (uint32_t)q - row_ptr[i];
        const uint32_t hs = l / LH, jl = l - hs * LH, lane = row * qq + hs;
        const uint64_t off = byte0[c] + (uint64_t)j * ec_task_bytes(L, nbt) + ec_hdr_bytes(L) +
                             4ull * (((uint64_t)b * LH + jl) * 32u + lane);
        *reinterpret_cast<float *>(blob + off) = nzval[a_csc[q]];
    }
}

// ---- the rest rows as a CSC of their own
__global__ void k_ec_keep(const uint32_t *__restrict__ rowval, int64_t nnz, const uint32_t *__restrict__ rest_flag, uint32_t *keep) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e <= nnz; e += (int64_t)gridDim.x * blockDim.x)
        keep[e] = (e < nnz && rest_flag[rowval[e] - 1u]) ? 1u : 0u;
}
__global__ void k_ec_emit(const uint32_t *__restrict__ rowval, const float *__restrict__ nzval, int64_t nnz,
                          const uint32_t *__restrict__ keep, const uint32_t *__restrict__ kept_before,
                          const uint32_t *__restrict__ rest_id, uint32_t *rowval_out, float *nzval_out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
        if (keep[e]) {
            rowval_out[kept_before[e]] = rest_id[rowval[e] - 1u] + 1u;
            nzval_out[kept_before[e]] = nzval[e];
        }
}
__global__ void k_ec_colptr(const uint32_t *__restrict__ colptr, int64_t n, const uint32_t *__restrict__ kept_before, uint32_t *out) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j <= n; j += (int64_t)gridDim.x * blockDim.x)
        out[j] = kept_before[colptr[j] - 1u] + 1u;
}
__global__ void k_ec_rest_rows(int64_t m, const uint32_t *__restrict__ rest_flag, const uint32_t *__restrict__ rest_id,
                               const int64_t *__restrict__ ks, uint32_t *rest_row, int64_t *ks_out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        if (rest_flag[i]) {
            rest_row[rest_id[i]] = (uint32_t)i;
            if (ks_out) ks_out[rest_id[i]] = ks[i];
        }
}

__global__ void k_ec_count_keys(const uint32_t *__restrict__ keys, uint32_t count, uint32_t *cnt) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) atomicAdd(&cnt[keys[i]], 1u);
}
__global__ void k_ec_iota(uint32_t *v, uint32_t count) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) v[i] = i;
}

int ec_bits_for(uint64_t maxval) {
    int b = 1;
    while (b < 64 && (maxval >> b) != 0) ++b;
    return b;
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
    void release_now(void *p) {
        for (auto &q : ptrs)
            if (q == p) {
                polee::dfree(p);
                q = nullptr;
            }
    }
};

struct EcPhaseTimer {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t0;
    explicit EcPhaseTimer(cudaStream_t s) : on(getenv("POLEE_SETUP_TIMING") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[polee setup] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

}  // namespace

void release_ec(polee_handle *h) {
    polee::dfree(h->ec_blob); polee::dfree(h->ec_desc); polee::dfree(h->ec_row_of_slot); polee::dfree(h->ec_slot_weight);
    polee::dfree(h->ec_units); polee::dfree(h->ec_multi); polee::dfree(h->rest_row);
    h->ec_blob = nullptr; h->ec_desc = nullptr; h->ec_row_of_slot = nullptr; h->ec_slot_weight = nullptr;
    h->ec_units = nullptr; h->ec_multi = nullptr; h->rest_row = nullptr;
    h->ec_tasks = 0; h->ec_rows = h->ec_nnz = h->ec_slots = h->ec_classes = 0; h->ec_blob_bytes = 0; h->ec_parts = 0;
    h->ec_nunits = h->ec_nmulti = h->ec_nlvl2 = 0;
}

#define CK(expr) POLEE_CUDA_CHECK(h, expr)

int setup_ec_from_device_csc(polee_handle *h, int64_t m, int64_t n, int64_t nnz, const uint32_t *d_colptr,
                             const uint32_t *d_rowval, const float *d_nzval, const int64_t *d_ks,
                             cudaEvent_t vals_ready_or_null, EcRest *rest) {
    release_ec(h);
    rest->m = m; rest->nnz = nnz;  // until proven otherwise everything is "rest"
    if (nnz < 1 || m < 1) return POLEE_OK;
    if (m > (int64_t)INT32_MAX || nnz > (int64_t)INT32_MAX) return POLEE_OK;  // CUB item counts are int here
    cudaStream_t st = h->stream;
    EcPhaseTimer pt(st);
    const int TPB = 256;
    auto grid_for = [&](int64_t work) { return (int)std::max<int64_t>(1, std::min<int64_t>((work + TPB - 1) / TPB, (int64_t)h->num_sms * 32)); };
    Scratch sc;
    size_t tmp_bytes = 0, need = 0;
    void *d_tmp = nullptr;
    auto ensure_tmp = [&](size_t bytes) -> cudaError_t {
        if (bytes <= tmp_bytes) return cudaSuccess;
        tmp_bytes = bytes + bytes / 8;
        return sc.alloc((char **)&d_tmp, tmp_bytes);
    };
    uint32_t min_rows = EC_MIN_ROWS_DEFAULT;
    const char *min_rows_env = getenv("POLEE_EC_MIN_ROWS");
    if (min_rows_env) min_rows = (uint32_t)std::max(1, atoi(min_rows_env));

    // ---- 1. CSR order
    uint32_t *row_len, *row_ptr, *col_of, *key_row, *val_e, *row_sorted, *a_csc, *col_csr;
    int *d_bad;
    CK(sc.alloc(&row_len, m + 1)); CK(sc.alloc(&row_ptr, m + 1)); CK(sc.alloc(&col_of, nnz)); CK(sc.alloc(&key_row, nnz));
    CK(sc.alloc(&val_e, nnz)); CK(sc.alloc(&row_sorted, nnz)); CK(sc.alloc(&a_csc, nnz)); CK(sc.alloc(&col_csr, nnz));
    CK(sc.alloc(&d_bad, 1));
    CK(cudaMemsetAsync(row_len, 0, sizeof(uint32_t) * (m + 1), st));
    CK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    k_ec_count<<<grid_for(nnz), TPB, 0, st>>>(d_rowval, nnz, row_len, m, d_bad);
    k_ec_expand<<<grid_for(nnz), TPB, 0, st>>>(d_colptr, n, d_rowval, nnz, col_of, key_row, val_e);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, key_row, row_sorted, val_e, a_csc, (int)nnz, 0, 32, st));
    CK(ensure_tmp(need));
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, key_row, row_sorted, val_e, a_csc, (int)nnz, 0, ec_bits_for((uint64_t)m), st));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, row_len, row_ptr, (int)(m + 1), st));
    CK(ensure_tmp(need));
    CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, row_len, row_ptr, (int)(m + 1), st));
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    pt.mark("ec: CSR order");

    // ---- 2. classes
    uint64_t *hash, *hash_s;
    uint32_t *row_id, *row_s, *is_new, *gid_incl, *rpos;
    CK(sc.alloc(&hash, m)); CK(sc.alloc(&hash_s, m)); CK(sc.alloc(&row_id, m)); CK(sc.alloc(&row_s, m));
    CK(sc.alloc(&is_new, m)); CK(sc.alloc(&gid_incl, m)); CK(sc.alloc(&rpos, m));
    k_ec_hash<<<grid_for(m), TPB, 0, st>>>(m, row_ptr, a_csc, col_of, col_csr, hash, row_id);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, hash, hash_s, row_id, row_s, (int)m, 0, 64, st));
    CK(ensure_tmp(need));
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, hash, hash_s, row_id, row_s, (int)m, 0, 64, st));
    k_ec_new<<<grid_for(m), TPB, 0, st>>>(m, hash_s, row_s, row_ptr, col_csr, is_new, rpos);
    CK(cub::DeviceScan::InclusiveSum(nullptr, need, is_new, gid_incl, (int)m, st));
    CK(ensure_tmp(need));
    CK(cub::DeviceScan::InclusiveSum(d_tmp, need, is_new, gid_incl, (int)m, st));
    uint32_t n_classes = 0;
    CK(cudaMemcpyAsync(&n_classes, gid_incl + (m - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (bad) return h->fail(POLEE_EINVAL, "set_matrix: rowval out of range 1..m");
    pt.mark("ec: classes");

    uint32_t *cstart, *leader, *gid_iota, *leader_s, *cls_s, *order_of_class, *cls_L, *cls_nt, *cls_nb, *cls_parts, *cls_rows;
    uint32_t *task0, *blk0, *part0, *rows0;
    uint64_t *cls_bytes, *byte0, *cls_ent, *ent0;
    const size_t nc1 = (size_t)n_classes + 1;
    CK(sc.alloc(&cstart, nc1)); CK(sc.alloc(&leader, nc1)); CK(sc.alloc(&gid_iota, nc1)); CK(sc.alloc(&leader_s, nc1));
    CK(sc.alloc(&cls_s, nc1)); CK(sc.alloc(&order_of_class, nc1)); CK(sc.alloc(&cls_L, nc1)); CK(sc.alloc(&cls_nt, nc1));
    CK(sc.alloc(&cls_nb, nc1)); CK(sc.alloc(&cls_parts, nc1)); CK(sc.alloc(&cls_rows, nc1)); CK(sc.alloc(&task0, nc1));
    CK(sc.alloc(&blk0, nc1)); CK(sc.alloc(&part0, nc1)); CK(sc.alloc(&rows0, nc1)); CK(sc.alloc(&cls_bytes, nc1));
    CK(sc.alloc(&byte0, nc1)); CK(sc.alloc(&cls_ent, nc1)); CK(sc.alloc(&ent0, nc1));
    k_ec_class_start<<<grid_for(m), TPB, 0, st>>>(m, is_new, gid_incl, row_s, n_classes, cstart, leader, gid_iota);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, leader, leader_s, gid_iota, cls_s, (int)n_classes, 0, 32, st));
    CK(ensure_tmp(need));
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, leader, leader_s, gid_iota, cls_s, (int)n_classes, 0, ec_bits_for((uint64_t)m), st));
    uint32_t n_tasks = 0, n_blocks = 0, n_parts = 0, ec_rows = 0;
    uint64_t total_bytes = 0, ec_ent = 0;
    auto build_table = [&](uint32_t min_r) -> int {
        k_ec_class_meta<<<grid_for(nc1), TPB, 0, st>>>(n_classes, cls_s, leader_s, cstart, row_len, min_r, order_of_class, cls_L,
                                                        cls_nt, cls_nb, cls_bytes, cls_parts, cls_rows, cls_ent);
        size_t need2 = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, need, cls_nt, task0, (int)nc1, st));
        CK(cub::DeviceScan::ExclusiveSum(nullptr, need2, cls_bytes, byte0, (int)nc1, st));
        CK(ensure_tmp(std::max(need, need2)));
        CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, cls_nt, task0, (int)nc1, st));
        CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, cls_nb, blk0, (int)nc1, st));
        CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, cls_parts, part0, (int)nc1, st));
        CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, cls_rows, rows0, (int)nc1, st));
        CK(cub::DeviceScan::ExclusiveSum(d_tmp, need2, cls_bytes, byte0, (int)nc1, st));
        CK(cub::DeviceScan::ExclusiveSum(d_tmp, need2, cls_ent, ent0, (int)nc1, st));
        CK(cudaMemcpyAsync(&n_tasks, task0 + n_classes, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&n_blocks, blk0 + n_classes, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&n_parts, part0 + n_classes, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&ec_rows, rows0 + n_classes, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&total_bytes, byte0 + n_classes, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&ec_ent, ent0 + n_classes, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return POLEE_OK;
    };
    // Small classes pay for their padding (a block is 32 rows whatever the class holds), but taking EVERY class spares
    // the step the general kernels altogether (their launches alone cost as much as ~0.5 GB of streaming): take all
    // classes when the padded layout stays below the 8 bytes per entry ONE pass of the general layouts streams, else
    // only classes of >= EC_MIN_ROWS_DEFAULT rows.  POLEE_EC_MIN_ROWS overrides.
    {
        int rc = build_table(min_rows_env ? min_rows : 1u);
        if (rc) return rc;
        if (!min_rows_env && (double)total_bytes > 8.0 * (double)ec_ent) {
            if ((rc = build_table(min_rows))) return rc;
        }
    }
    pt.mark("ec: class table");
    if (getenv("POLEE_SETUP_TIMING"))
        fprintf(stderr, "[polee setup] ec: %u classes; %u rows (%.1f %%) / %llu entries (%.1f %%) in %u tasks, %u blocks, %.1f MB "
                        "(%.2f B/entry), %u partials\n", n_classes, ec_rows, 100.0 * ec_rows / m, (unsigned long long)ec_ent,
                100.0 * ec_ent / nnz, n_tasks, n_blocks, total_bytes / 1e6, ec_ent ? (double)total_bytes / ec_ent : 0.0, n_parts);
    if (n_tasks == 0) return POLEE_OK;
    if ((uint64_t)n_blocks * 32ull >= 0xFFFFFFFFull) return POLEE_OK;  // row slots are 32-bit

    // ---- 3. blobs
    const uint64_t n_slots = (uint64_t)n_blocks * 32ull;
    CK(polee::dmalloc((void **)&h->ec_blob, total_bytes + 16));
    CK(polee::dmalloc((void **)&h->ec_desc, sizeof(EcTaskDesc) * n_tasks));
    CK(polee::dmalloc((void **)&h->ec_row_of_slot, sizeof(uint32_t) * n_slots));
    if (d_ks) CK(polee::dmalloc((void **)&h->ec_slot_weight, sizeof(float) * n_slots));
    CK(cudaMemsetAsync(h->ec_blob, 0, total_bytes + 16, st));
    CK(cudaMemsetAsync(h->ec_row_of_slot, 0xFF, sizeof(uint32_t) * n_slots, st));
    if (d_ks) CK(cudaMemsetAsync(h->ec_slot_weight, 0, sizeof(float) * n_slots, st));
    uint32_t *part_col, *part_col_s, *part_task, *pid_iota, *plist, *task_part0, *col_cnt, *slot_of_row, *rest_flag, *rest_id;
    CK(sc.alloc(&part_col, n_parts)); CK(sc.alloc(&part_col_s, n_parts)); CK(sc.alloc(&part_task, n_parts));
    CK(sc.alloc(&pid_iota, n_parts)); CK(sc.alloc(&plist, n_parts)); CK(sc.alloc(&task_part0, n_tasks)); CK(sc.alloc(&col_cnt, n));
    CK(sc.alloc(&slot_of_row, m)); CK(sc.alloc(&rest_flag, m + 1)); CK(sc.alloc(&rest_id, m + 1));
    k_ec_task_init<<<grid_for(n_tasks), TPB, 0, st>>>(n_tasks, n_classes, task0, blk0, byte0, part0, cls_L, cls_nb, cls_rows,
                                                       leader_s, row_ptr, col_csr, h->ec_desc, h->ec_blob, part_col, part_task);
    k_ec_rows<<<grid_for(m + 1), TPB, 0, st>>>(m, row_s, gid_incl, order_of_class, cstart, cls_L, blk0, d_ks, h->ec_row_of_slot,
                                                h->ec_slot_weight, slot_of_row, rest_flag);
    if (vals_ready_or_null) CK(cudaStreamWaitEvent(st, vals_ready_or_null, 0));
    k_ec_fill<<<grid_for(nnz), TPB, 0, st>>>(nnz, row_sorted, a_csc, row_ptr, slot_of_row, rpos, gid_incl, order_of_class, cls_L,
                                              blk0, byte0, d_nzval, h->ec_blob);
    pt.mark("ec: pack blobs");

    // ---- 4. partial slots by column + the second-stage work list
    CK(cudaMemsetAsync(col_cnt, 0, sizeof(uint32_t) * n, st));
    k_ec_iota<<<grid_for(n_parts), TPB, 0, st>>>(pid_iota, n_parts);
    k_ec_task_part0<<<grid_for(n_parts), TPB, 0, st>>>(n_parts, part_task, task_part0);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, part_col, part_col_s, pid_iota, plist, (int)n_parts, 0, 32, st));
    CK(ensure_tmp(need));
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, part_col, part_col_s, pid_iota, plist, (int)n_parts, 0, ec_bits_for((uint64_t)n), st));
    k_ec_count_keys<<<grid_for(n_parts), TPB, 0, st>>>(part_col, n_parts, col_cnt);
    k_ec_dest<<<grid_for(n_parts), TPB, 0, st>>>(n_parts, plist, part_task, task_part0, h->ec_desc, h->ec_blob);
    std::vector<uint32_t> cnt(n);
    CK(cudaMemcpyAsync(cnt.data(), col_cnt, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    {   // processing order (see k_ec_task_key); stable, so tasks of one kind keep the matrix order
        uint32_t *tkey, *tkey_s;
        EcTaskDesc *desc_sorted = nullptr;
        CK(sc.alloc(&tkey, n_tasks)); CK(sc.alloc(&tkey_s, n_tasks));
        CK(polee::dmalloc((void **)&desc_sorted, sizeof(EcTaskDesc) * n_tasks));
        k_ec_task_key<<<grid_for(n_tasks), TPB, 0, st>>>(n_tasks, h->ec_desc, h->ec_blob, tkey);
        cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, need, tkey, tkey_s, h->ec_desc, desc_sorted, (int)n_tasks, 0, 5, st);
        if (e == cudaSuccess) e = ensure_tmp(need);
        if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(d_tmp, need, tkey, tkey_s, h->ec_desc, desc_sorted, (int)n_tasks, 0, 5, st);
        if (e != cudaSuccess) {
            polee::dfree(desc_sorted);
            CK(e);
        }
        uint32_t *kcnt;
        CK(sc.alloc(&kcnt, 32));
        CK(cudaMemsetAsync(kcnt, 0, sizeof(uint32_t) * 32, st));
        k_ec_count_keys<<<grid_for(n_tasks), TPB, 0, st>>>(tkey, n_tasks, kcnt);
        uint32_t hk[32];
        CK(cudaMemcpyAsync(hk, kcnt, sizeof(hk), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        polee::dfree(h->ec_desc);
        h->ec_desc = desc_sorted;
        int run_end = 0;
        for (int lh = 8; lh >= 1; --lh) {  // key = 31 - ((LH - 1) * 4 + lq)
            for (int lq = 0; lq < 4; ++lq) run_end += (int)hk[31 - ((lh - 1) * 4 + lq)];
            h->ec_kind_end[8 - lh] = run_end;
        }
    }

    // ---- 5. the rest rows
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, rest_flag, rest_id, (int)(m + 1), st));
    CK(ensure_tmp(need));
    CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, rest_flag, rest_id, (int)(m + 1), st));
    uint32_t m_rest = 0;
    CK(cudaMemcpyAsync(&m_rest, rest_id + m, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    rest->m = m_rest;
    rest->nnz = nnz - (int64_t)ec_ent;
    if (m_rest > 0) {
        uint32_t *keep, *kept_before;
        CK(sc.alloc(&keep, nnz + 1)); CK(sc.alloc(&kept_before, nnz + 1));
        k_ec_keep<<<grid_for(nnz + 1), TPB, 0, st>>>(d_rowval, nnz, rest_flag, keep);
        CK(cub::DeviceScan::ExclusiveSum(nullptr, need, keep, kept_before, (int)(nnz + 1), st));
        CK(ensure_tmp(need));
        CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, keep, kept_before, (int)(nnz + 1), st));
        CK(polee::dmalloc((void **)&rest->colptr, sizeof(uint32_t) * (n + 1)));
        CK(polee::dmalloc((void **)&rest->rowval, sizeof(uint32_t) * std::max<int64_t>(rest->nnz, 1)));
        CK(polee::dmalloc((void **)&rest->nzval, sizeof(float) * std::max<int64_t>(rest->nnz, 1)));
        CK(polee::dmalloc((void **)&h->rest_row, sizeof(uint32_t) * m_rest));
        if (d_ks) CK(polee::dmalloc((void **)&rest->ks, sizeof(int64_t) * m_rest));
        k_ec_emit<<<grid_for(nnz), TPB, 0, st>>>(d_rowval, d_nzval, nnz, keep, kept_before, rest_id, rest->rowval, rest->nzval);
        k_ec_colptr<<<grid_for(n + 1), TPB, 0, st>>>(d_colptr, n, kept_before, rest->colptr);
        k_ec_rest_rows<<<grid_for(m), TPB, 0, st>>>(m, rest_flag, rest_id, d_ks, h->rest_row, rest->ks);
        CK(cudaStreamSynchronize(st));
    }
    pt.mark("ec: partial slots + rest rows");

    std::vector<FusedUnit> units;
    std::vector<FusedMulti> multi;
    units.reserve((size_t)n + n_parts / FT_UNIT);
    uint32_t pos = 0, lvl2 = 0;
    for (int64_t j = 0; j < n; ++j) {
        const uint32_t c = cnt[j], nu = c == 0 ? 1u : (c + FT_UNIT - 1) / FT_UNIT;
        if (nu == 1) {
            units.push_back(FusedUnit{(uint32_t)j, pos, pos + c, -1});
        } else {
            multi.push_back(FusedMulti{(uint32_t)j, lvl2, nu, 0u});
            for (uint32_t u = 0; u < nu; ++u)
                units.push_back(FusedUnit{(uint32_t)j, pos + u * FT_UNIT, pos + std::min<uint32_t>(c, (u + 1) * FT_UNIT), (int32_t)lvl2++});
        }
        pos += c;
    }
    CK(polee::dmalloc((void **)&h->ec_units, sizeof(FusedUnit) * std::max<size_t>(units.size(), 1)));
    CK(polee::dmalloc((void **)&h->ec_multi, sizeof(FusedMulti) * std::max<size_t>(multi.size(), 1)));
    CK(cudaMemcpyAsync(h->ec_units, units.data(), sizeof(FusedUnit) * units.size(), cudaMemcpyHostToDevice, st));
    if (!multi.empty())
        CK(cudaMemcpyAsync(h->ec_multi, multi.data(), sizeof(FusedMulti) * multi.size(), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    h->ec_tasks = (int)n_tasks;
    h->ec_rows = ec_rows; h->ec_nnz = (int64_t)ec_ent; h->ec_slots = (int64_t)n_slots; h->ec_classes = n_classes;
    h->ec_blob_bytes = total_bytes;
    h->ec_parts = n_parts;
    h->ec_nunits = (int)units.size();
    h->ec_nmulti = (int)multi.size();
    h->ec_nlvl2 = (int)lvl2;
    pt.mark("ec: second-stage list");
    return POLEE_OK;
}

}  // namespace polee
