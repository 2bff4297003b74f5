// matrix_setup.cu -- one-time (per fit) conversion of RNASeqSample.X, handed over exactly as Julia
// stores it (SparseMatrixCSC{Float32,UInt32}: 1-based colptr/rowval, src/rnaseq_sample.jl:11,499), into
// the two HBM layouts the per-step kernels stream.  Replaces `Xt = SparseMatrixCSC(transpose(X))`
// (src/likelihood-approximation.jl:407), which the reference does on the host.
//
//  K1 layout ("SELL slabs"): rows (fragments) are stably sorted by their entry count, longest class
//  first, and renumbered; class L is stored as a dense L x stride slab, t-major, so that lane r of a
//  warp reads entry t of row r with a perfectly coalesced access and all 32 rows of a warp have the
//  same trip count.  Inside a row the entries keep ascending transcript order -- the order in which
//  the reference accumulates (src/sparse.jl:14-18).  No row pointer array is streamed at all.
//
//  K2 layout: the CSC arrays with row ids replaced by the permuted ids and each column re-sorted by
//  permuted row id, so that consecutive lanes of a warp gather neighbouring rows of w.  Columns are cut
//  into segments of <= COL_SEG entries (one warp each).
//
// The sorts use CUB's device radix sort (a library call, but setup only -- not on the per-step path).
#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace polee {

namespace {

__global__ void k_count_rows(const uint32_t *__restrict__ rowval, int64_t nnz, uint32_t *row_len, int64_t m,
                             int *bad) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        uint32_t r = rowval[e] - 1u;
        if (r >= (uint64_t)m)
            *bad = 1;
        else
            atomicAdd(&row_len[r], 1u);
    }
}

__global__ void k_make_len_keys(const uint32_t *__restrict__ row_len, int64_t m, uint32_t lmax, uint32_t *keys,
                                uint32_t *vals) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = lmax - row_len[i];
        vals[i] = (uint32_t)i;
    }
}

// position in the length-sorted order -> permuted (class-padded) row id
__global__ void k_assign_perm(const uint32_t *__restrict__ sorted_rows, const uint32_t *__restrict__ row_len,
                              int64_t m, const uint32_t *__restrict__ cls_row_off_pad,
                              const uint32_t *__restrict__ cls_row_off_unpad, uint32_t *row_perm, uint32_t *len_perm) {
    for (int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pos < m;
         pos += (int64_t)gridDim.x * blockDim.x) {
        uint32_t i = sorted_rows[pos];
        uint32_t L = row_len[i];
        uint32_t rp = cls_row_off_pad[L] + (uint32_t)(pos - cls_row_off_unpad[L]);
        row_perm[i] = rp;
        len_perm[rp] = L;
    }
}

__global__ void k_entry_keys(const uint32_t *__restrict__ colptr, int64_t n, const uint32_t *__restrict__ rowval,
                             int64_t nnz, const uint32_t *__restrict__ row_perm, uint32_t *col_of, uint32_t *key_r,
                             uint32_t *val_e) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        // largest j with colptr[j] - 1 <= e
        int64_t lo = 0, hi = n;  // invariant: colptr[lo]-1 <= e < colptr[hi]-1
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if ((int64_t)colptr[mid] - 1 <= e)
                lo = mid;
            else
                hi = mid;
        }
        col_of[e] = (uint32_t)lo;
        key_r[e] = row_perm[rowval[e] - 1u];
        val_e[e] = (uint32_t)e;
    }
}

__global__ void k_row_starts(const uint32_t *__restrict__ sorted_r, int64_t nnz, uint32_t *row_start) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
        if (q == 0 || sorted_r[q] != sorted_r[q - 1]) row_start[sorted_r[q]] = (uint32_t)q;
    }
}

__global__ void k_fill_sell(const uint32_t *__restrict__ sorted_r, const uint32_t *__restrict__ sorted_e, int64_t nnz,
                            const uint32_t *__restrict__ row_start, const uint32_t *__restrict__ len_perm,
                            const uint32_t *__restrict__ cls_row_off_pad, const uint64_t *__restrict__ cls_slab_off,
                            const uint32_t *__restrict__ cls_stride, const uint32_t *__restrict__ col_of,
                            const float *__restrict__ nzval, uint32_t *sell_idx, float *sell_val, uint32_t *key_c,
                            uint32_t *val_q) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
        uint32_t rp = sorted_r[q], e = sorted_e[q];
        uint32_t L = len_perm[rp];
        uint32_t t = (uint32_t)q - row_start[rp];
        uint64_t pos = cls_slab_off[L] + (uint64_t)t * cls_stride[L] + (rp - cls_row_off_pad[L]);
        uint32_t c = col_of[e];
        sell_idx[pos] = c;
        sell_val[pos] = nzval[e];
        key_c[q] = c;
        val_q[q] = (uint32_t)q;
    }
}

__global__ void k_fill_csc(const uint32_t *__restrict__ order_q, int64_t nnz, const uint32_t *__restrict__ sorted_r,
                           const uint32_t *__restrict__ sorted_e, const float *__restrict__ nzval, uint32_t *csc_row,
                           float *csc_val) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        uint32_t q = order_q[p];
        csc_row[p] = sorted_r[q];
        csc_val[p] = nzval[sorted_e[q]];
    }
}

__global__ void k_row_weights(const int64_t *__restrict__ ks, int64_t m, const uint32_t *__restrict__ row_perm,
                              float *row_weight) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        row_weight[row_perm[i]] = (float)ks[i];
}


// ------------------------------------------------------------------ fused row-tile layout (see common.cuh)
__global__ void k_tile_flags(const uint32_t *__restrict__ row_ptr, int64_t m, uint32_t tile_rows, uint32_t window,
                             uint32_t *flag) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = (i % tile_rows == 0 || row_ptr[i] / window != row_ptr[i - 1] / window) ? 1u : 0u;
}

// tile_incl[i] = 1 + tile of row i (inclusive scan of the flags)
__global__ void k_tile_row0(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ tile_incl, int64_t m,
                            uint32_t n_tiles, uint32_t *tile_row0) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        if (flag[i]) tile_row0[tile_incl[i] - 1u] = (uint32_t)i;
        if (i == m - 1) tile_row0[n_tiles] = (uint32_t)m;
    }
}

// column of every CSC entry + an order-independent hash of every row's column set
__global__ void k_col_of(const uint32_t *__restrict__ colptr, int64_t n, const uint32_t *__restrict__ rowval, int64_t nnz,
                         uint32_t *col_of, uint32_t *row_pattern) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n;  // largest j with colptr[j] - 1 <= e
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if ((int64_t)colptr[mid] - 1 <= e)
                lo = mid;
            else
                hi = mid;
        }
        col_of[e] = (uint32_t)lo;
        uint32_t hsh = (uint32_t)lo * 0x9E3779B1u;
        hsh ^= hsh >> 15;
        hsh *= 0x85EBCA77u;
        atomicAdd(&row_pattern[rowval[e] - 1u], hsh ^ (hsh >> 13));
    }
}

// rows of a tile longest first, rows with the same column set next to each other (the lanes of a warp then read the
// same x from shared memory -- a broadcast -- and a column's entries hit consecutive rows of w):
// key = tile << 44 | (0xFFFF - len) << 28 | pattern hash (28 bits)
__global__ void k_row_sort_keys(const uint32_t *__restrict__ tile_incl, const uint32_t *__restrict__ row_len,
                                const uint32_t *__restrict__ row_pattern, int64_t m, uint64_t *keys, uint32_t *vals) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = ((uint64_t)(tile_incl[i] - 1u) << 44) | ((uint64_t)(0xFFFFu - min(row_len[i], 0xFFFFu)) << 28) |
                  (uint64_t)(row_pattern[i] & 0x0FFFFFFFu);
        vals[i] = (uint32_t)i;
    }
}

__global__ void k_row_positions(const uint32_t *__restrict__ row_of_pos, const uint32_t *__restrict__ row_len, int64_t m,
                                uint32_t *rpos, uint32_t *len_sorted) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= m; p += (int64_t)gridDim.x * blockDim.x) {
        if (p == m) {
            len_sorted[p] = 0;
            continue;
        }
        const uint32_t i = row_of_pos[p];
        rpos[i] = (uint32_t)p;
        len_sorted[p] = row_len[i];
    }
}

__global__ void k_csc_expand(const uint32_t *__restrict__ rowval, int64_t nnz, const uint32_t *__restrict__ rpos,
                             uint32_t *key_row, uint32_t *val_e) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        key_row[e] = rpos[rowval[e] - 1u];
        val_e[e] = (uint32_t)e;
    }
}

// second order: the row-major list re-sorted (stably) by (tile, column) -> (tile, column, row position)
__global__ void k_b_keys(const uint32_t *__restrict__ posA, const uint32_t *__restrict__ a_csc, int64_t nnz,
                         const uint32_t *__restrict__ row_of_pos, const uint32_t *__restrict__ tile_incl,
                         const uint32_t *__restrict__ col_of, uint64_t *keys) {
    for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nnz; a += (int64_t)gridDim.x * blockDim.x)
        keys[a] = ((uint64_t)(tile_incl[row_of_pos[posA[a]]] - 1u) << 32) | (uint64_t)col_of[a_csc[a]];
}

__global__ void k_b_tiles(const uint64_t *__restrict__ keys_sorted, int64_t nnz, uint32_t *tileB) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x)
        tileB[q] = (uint32_t)(keys_sorted[q] >> 32);
}

__global__ void k_inverse_perm(const uint32_t *__restrict__ a_csc, int64_t nnz, uint32_t *apos_of_csc) {
    for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nnz; a += (int64_t)gridDim.x * blockDim.x)
        apos_of_csc[a_csc[a]] = (uint32_t)a;
}

// column-major (within tile) order: (tile, column) run starts (+ a zero sentinel at nnz for the exclusive scan)
__global__ void k_col_starts(const uint32_t *__restrict__ tileB, const uint32_t *__restrict__ b_csc, int64_t nnz,
                             const uint32_t *__restrict__ col_of, uint32_t *colstart) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q <= nnz; q += (int64_t)gridDim.x * blockDim.x) {
        if (q == nnz) {
            colstart[q] = 0;
            continue;
        }
        colstart[q] = (q == 0 || tileB[q] != tileB[q - 1] || col_of[b_csc[q]] != col_of[b_csc[q - 1]]) ? 1u : 0u;
    }
}

// run_pos[pid] = first position of (tile, column) run pid
__global__ void k_run_pos(const uint32_t *__restrict__ colstart, const uint32_t *__restrict__ cols_before, int64_t nnz,
                          uint32_t *run_pos) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x)
        if (colstart[q]) run_pos[cols_before[q]] = (uint32_t)q;
}

// segment starts: every FT_SEG entries of a run
__global__ void k_seg_starts(const uint32_t *__restrict__ colstart, const uint32_t *__restrict__ cols_before,
                             const uint32_t *__restrict__ run_pos, int64_t nnz, uint32_t *segstart) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q <= nnz; q += (int64_t)gridDim.x * blockDim.x) {
        if (q == nnz) {
            segstart[q] = 0;
            continue;
        }
        const uint32_t pid = cols_before[q] + colstart[q] - 1u;
        segstart[q] = (((uint32_t)q - run_pos[pid]) % FT_SEG == 0u) ? 1u : 0u;
    }
}

__global__ void k_tile_meta(uint32_t n_tiles, const uint32_t *__restrict__ tile_row0, const uint32_t *__restrict__ row_ptr,
                            const uint32_t *__restrict__ len_sorted, const uint32_t *__restrict__ segs_before,
                            const uint32_t *__restrict__ cols_before, FusedHdr *hdrs, uint32_t *col0, uint64_t *blob_bytes,
                            uint32_t *maxima /* E, C, rows, bytes, S, EA */) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
        const uint32_t row0 = tile_row0[t], row1 = tile_row0[t + 1];
        const uint32_t a0 = row_ptr[row0], a1 = row_ptr[row1];
        FusedHdr hd;
        hd.rows = row1 - row0;
        hd.E = a1 - a0;
        hd.C = cols_before[a1] - cols_before[a0];
        hd.S = segs_before[a1] - segs_before[a0];
        hd.row0 = row0;
        hd.part0 = segs_before[a0];
        hd.G = (hd.rows + 31u) / 32u;
        uint32_t ea = 0;
        for (uint32_t g = 0; g < hd.G; ++g) ea += 32u * len_sorted[row0 + 32u * g];  // rows are sorted longest first
        hd.EA = ea;
        hdrs[t] = hd;
        col0[t] = cols_before[a0];
        const uint32_t bytes = blob_layout(hd).bytes;
        blob_bytes[t] = bytes;
        atomicMax(&maxima[0], hd.E);
        atomicMax(&maxima[1], hd.C);
        atomicMax(&maxima[2], hd.rows);
        atomicMax(&maxima[3], bytes);
        atomicMax(&maxima[4], hd.S);
        atomicMax(&maxima[5], hd.EA);
    }
}

// header + group table of every blob
__global__ void k_tile_desc(uint32_t n_tiles, const FusedHdr *__restrict__ hdrs, const uint64_t *__restrict__ blob_off,
                            const uint64_t *__restrict__ blob_bytes, const uint32_t *__restrict__ len_sorted,
                            FusedTileDesc *desc, unsigned char *blob) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
        const FusedHdr hd = hdrs[t];
        desc[t] = FusedTileDesc{blob_off[t], (uint32_t)blob_bytes[t], 0u};
        unsigned char *b = blob + blob_off[t];
        *reinterpret_cast<FusedHdr *>(b) = hd;
        uint32_t *ginfo = reinterpret_cast<uint32_t *>(b + blob_layout(hd).ginfo);
        uint32_t base = 0;
        for (uint32_t g = 0; g < hd.G; ++g) {
            const uint32_t L = len_sorted[hd.row0 + 32u * g];
            ginfo[g] = base | (L << 16);
            base += 32u * L;
        }
    }
}

// row-side slab position of the t-th entry of the row at tile-local position rr
__device__ __forceinline__ uint32_t slab_index(const unsigned char *b, const BlobLayout &L, uint32_t rr, uint32_t t) {
    const uint32_t gi = reinterpret_cast<const uint32_t *>(b + L.ginfo)[rr >> 5];
    return (gi & 0xffffu) + t * 32u + (rr & 31u);
}

// a = position in the (row position, column) order
__global__ void k_pack_a(int64_t nnz, const uint32_t *__restrict__ posA, const uint32_t *__restrict__ a_csc,
                         const uint32_t *__restrict__ row_of_pos, const uint32_t *__restrict__ tile_incl,
                         const uint32_t *__restrict__ row_ptr, const FusedHdr *__restrict__ hdrs,
                         const uint64_t *__restrict__ blob_off, const float *__restrict__ nzval, unsigned char *blob) {
    for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nnz; a += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t p = posA[a];
        const uint32_t t = tile_incl[row_of_pos[p]] - 1u;
        const FusedHdr hd = hdrs[t];
        const BlobLayout L = blob_layout(hd);
        unsigned char *b = blob + blob_off[t];
        const uint32_t idx = slab_index(b, L, p - hd.row0, (uint32_t)a - row_ptr[p]);
        reinterpret_cast<float *>(b + L.valA)[idx] = nzval[a_csc[a]];
    }
}

// q = position in the (tile, column, row) order
__global__ void k_pack_b(int64_t nnz, const uint32_t *__restrict__ tileB, const uint32_t *__restrict__ b_csc,
                         const uint32_t *__restrict__ apos_of_csc, const uint32_t *__restrict__ posA,
                         const uint32_t *__restrict__ col_of, const uint32_t *__restrict__ segstart,
                         const uint32_t *__restrict__ colstart, const uint32_t *__restrict__ segs_before,
                         const uint32_t *__restrict__ cols_before, const uint32_t *__restrict__ row_ptr,
                         const FusedHdr *__restrict__ hdrs, const uint32_t *__restrict__ col0,
                         const uint64_t *__restrict__ blob_off, const float *__restrict__ nzval, unsigned char *blob,
                         uint32_t *part_col, uint32_t *part_tile) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t t = tileB[q];
        const FusedHdr hd = hdrs[t];
        const BlobLayout L = blob_layout(hd);
        unsigned char *b = blob + blob_off[t];
        const uint32_t a0 = row_ptr[hd.row0];
        const uint32_t qq = (uint32_t)q - a0;
        const uint32_t src = b_csc[q];
        const uint32_t apos = apos_of_csc[src];
        const uint32_t p = posA[apos];
        const uint32_t pid = cols_before[q] + colstart[q] - 1u;  // (tile, column) run containing q
        reinterpret_cast<float *>(b + L.valB)[qq] = nzval[src];
        reinterpret_cast<uint16_t *>(b + L.lrowB)[qq] = (uint16_t)(p - hd.row0);
        (b + L.lcolA)[slab_index(b, L, p - hd.row0, apos - row_ptr[p])] = (unsigned char)(pid - col0[t]);
        if (segstart[q]) {
            const uint32_t sid = segs_before[q];
            reinterpret_cast<uint16_t *>(b + L.segptr)[sid - hd.part0] = (uint16_t)qq;
            part_col[sid] = col_of[src];
            part_tile[sid] = t;
        }
        if (colstart[q]) reinterpret_cast<uint32_t *>(b + L.cols)[pid - col0[t]] = col_of[src];
        if (qq == hd.E - 1u) reinterpret_cast<uint16_t *>(b + L.segptr)[hd.S] = (uint16_t)hd.E;
    }
}

// the partial array is ordered by column: segment plist[i] writes slot i
__global__ void k_pack_dest(uint32_t n_parts, const uint32_t *__restrict__ plist, const uint32_t *__restrict__ part_tile,
                            const FusedHdr *__restrict__ hdrs, const uint64_t *__restrict__ blob_off, unsigned char *blob) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_parts; i += gridDim.x * blockDim.x) {
        const uint32_t sid = plist[i], t = part_tile[sid];
        const FusedHdr hd = hdrs[t];
        reinterpret_cast<uint32_t *>(blob + blob_off[t] + blob_layout(hd).dest)[sid - hd.part0] = i;
    }
}

__global__ void k_iota_u32(uint32_t *v, int64_t count) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = (uint32_t)i;
}

__global__ void k_count_keys(const uint32_t *__restrict__ keys, int64_t count, uint32_t *cnt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&cnt[keys[i]], 1u);
}

__global__ void k_ks_to_f32(const int64_t *__restrict__ ks, const uint32_t *__restrict__ row_of_pos, int64_t m, float *out) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m; p += (int64_t)gridDim.x * blockDim.x)
        out[p] = (float)ks[row_of_pos[p]];
}

int bits_for(uint64_t maxval) {
    int b = 1;
    while (b < 32 && (maxval >> b) != 0) ++b;
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
};

}  // namespace

void release_fused(polee_handle *h);

void release_matrix(polee_handle *h) {
    polee::dfree(h->sell_idx); polee::dfree(h->sell_val); polee::dfree(h->row_tiles); polee::dfree(h->row_perm);
    polee::dfree(h->row_weight); polee::dfree(h->csc_row); polee::dfree(h->csc_val); polee::dfree(h->segs); polee::dfree(h->multi);
    h->sell_idx = nullptr; h->sell_val = nullptr; h->row_tiles = nullptr; h->row_perm = nullptr;
    h->row_weight = nullptr; h->csc_row = nullptr; h->csc_val = nullptr; h->segs = nullptr; h->multi = nullptr;
    h->n_row_tiles = 0; h->n_segs = 0; h->n_multi = 0; h->n_slots = 0; h->m_pad = 0;
    release_fused(h);
    release_ec(h);
    h->gm = h->gnnz = 0;
    h->have_matrix = false;
}

#define CK(expr) POLEE_CUDA_CHECK(h, expr)

// POLEE_SETUP_TIMING=1 prints where the layout build spends its time (synchronises at every mark)
struct PhaseTimer {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(cudaStream_t s) : on(getenv("POLEE_SETUP_TIMING") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[polee setup] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

void release_fused(polee_handle *h) {
    polee::dfree(h->ft_blob); polee::dfree(h->ft_desc); polee::dfree(h->ft_units);
    polee::dfree(h->ft_multi); polee::dfree(h->ft_row_weight); polee::dfree(h->ft_row_of_pos);
    h->ft_row_of_pos = nullptr;
    h->ft_blob = nullptr; h->ft_desc = nullptr; h->ft_units = nullptr; h->ft_multi = nullptr;
    h->ft_row_weight = nullptr;
    h->fused = false;
    h->ft_tiles = 0; h->ft_parts = 0; h->ft_nunits = h->ft_nmulti = h->ft_nlvl2 = 0;
}

// Build the fused layout.  All passes are device passes (CUB sorts / scans + small kernels); the host only sees the
// per-column partial counts (n numbers) to cut the second-stage work list.
int setup_fused_from_device_csc(polee_handle *h, int64_t m, int64_t n, int64_t nnz, const uint32_t *d_colptr,
                                const uint32_t *d_rowval, const float *d_nzval, const int64_t *d_ks,
                                const std::vector<uint32_t> &colptr_host, cudaEvent_t vals_ready_or_null) {
    release_fused(h);
    cudaStream_t st = h->stream;
    PhaseTimer pt(st);
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

    // ---- row lengths, row pointers, tiles
    uint32_t *row_len, *row_ptr, *flag, *tile_incl, *d_lmax;
    int *d_bad;
    CK(sc.alloc(&row_len, m + 1)); CK(sc.alloc(&row_ptr, m + 1)); CK(sc.alloc(&flag, m)); CK(sc.alloc(&tile_incl, m));
    CK(sc.alloc(&d_lmax, 1)); CK(sc.alloc(&d_bad, 1));
    CK(cudaMemsetAsync(row_len, 0, sizeof(uint32_t) * (m + 1), st));
    CK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    if (nnz > 0) k_count_rows<<<grid_for(nnz), TPB, 0, st>>>(d_rowval, nnz, row_len, m, d_bad);
    CK(cub::DeviceReduce::Max(nullptr, need, row_len, d_lmax, (int)m, st));
    CK(ensure_tmp(need));
    CK(cub::DeviceReduce::Max(d_tmp, need, row_len, d_lmax, (int)m, st));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, row_len, row_ptr, (int)(m + 1), st));
    CK(ensure_tmp(need));
    CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, row_len, row_ptr, (int)(m + 1), st));
    uint32_t tile_rows = FT_ROWS, window = FT_ENTRY_WINDOW;
    if (const char *e = getenv("POLEE_FT_ROWS")) tile_rows = std::max(32, atoi(e)) / 32 * 32;
    if (const char *e = getenv("POLEE_FT_WINDOW")) window = (uint32_t)std::max(256, atoi(e));
    k_tile_flags<<<grid_for(m), TPB, 0, st>>>(row_ptr, m, tile_rows, window, flag);
    CK(cub::DeviceScan::InclusiveSum(nullptr, need, flag, tile_incl, (int)m, st));
    CK(ensure_tmp(need));
    CK(cub::DeviceScan::InclusiveSum(d_tmp, need, flag, tile_incl, (int)m, st));
    int bad = 0;
    uint32_t lmax = 0, n_tiles = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&lmax, d_lmax, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&n_tiles, tile_incl + (m - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (bad) return h->fail(POLEE_EINVAL, "set_matrix: rowval out of range 1..m");
    if ((uint64_t)lmax + FT_ENTRY_WINDOW > 32767ull) return POLEE_OK;  // a row too long for 16-bit tile offsets
    auto unsuitable = [&]() { release_fused(h); return (int)POLEE_OK; };
    pt.mark("fused: rows + tiles");
    uint32_t *tile_row0;
    CK(sc.alloc(&tile_row0, (size_t)n_tiles + 1));
    k_tile_row0<<<grid_for(m), TPB, 0, st>>>(flag, tile_incl, m, n_tiles, tile_row0);

    // ---- rows of every tile longest first; row pointers in that order
    uint64_t *rkey, *rkey_s;
    uint32_t *rval, *rpos, *len_sorted, *row_pattern, *col_of;
    CK(sc.alloc(&rkey, m)); CK(sc.alloc(&rkey_s, m)); CK(sc.alloc(&rval, m)); CK(sc.alloc(&rpos, m)); CK(sc.alloc(&len_sorted, m + 1));
    CK(sc.alloc(&row_pattern, m)); CK(sc.alloc(&col_of, nnz + 1));
    CK(polee::dmalloc((void **)&h->ft_row_of_pos, sizeof(uint32_t) * m));
    uint32_t *d_colptr_own = nullptr;
    if (!d_colptr) {
        CK(sc.alloc(&d_colptr_own, n + 1));
        CK(cudaMemcpyAsync(d_colptr_own, colptr_host.data(), 4 * (n + 1), cudaMemcpyHostToDevice, st));
        d_colptr = d_colptr_own;
    }
    CK(cudaMemsetAsync(row_pattern, 0, sizeof(uint32_t) * m, st));
    if (nnz > 0) k_col_of<<<grid_for(nnz), TPB, 0, st>>>(d_colptr, n, d_rowval, nnz, col_of, row_pattern);
    k_row_sort_keys<<<grid_for(m), TPB, 0, st>>>(tile_incl, row_len, row_pattern, m, rkey, rval);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, rkey, rkey_s, rval, h->ft_row_of_pos, (int)m, 0, 64, st));
    CK(ensure_tmp(need));
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, rkey, rkey_s, rval, h->ft_row_of_pos, (int)m, 0,
                                       44 + bits_for((uint64_t)n_tiles), st));
    k_row_positions<<<grid_for(m + 1), TPB, 0, st>>>(h->ft_row_of_pos, row_len, m, rpos, len_sorted);
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, len_sorted, row_ptr, (int)(m + 1), st));
    CK(ensure_tmp(need));
    CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, len_sorted, row_ptr, (int)(m + 1), st));

    // ---- the two entry orders: row-major (A) and column-major inside a tile (B)
    uint32_t *keyR, *valE, *posA, *a_csc, *tileB, *b_csc, *apos_of_csc;
    uint64_t *keyB, *keyB_s;
    CK(sc.alloc(&keyR, nnz)); CK(sc.alloc(&valE, nnz)); CK(sc.alloc(&posA, nnz)); CK(sc.alloc(&a_csc, nnz));
    CK(sc.alloc(&tileB, nnz)); CK(sc.alloc(&b_csc, nnz + 1)); CK(sc.alloc(&apos_of_csc, nnz));
    CK(sc.alloc(&keyB, nnz)); CK(sc.alloc(&keyB_s, nnz));
    if (nnz > 0) {
        k_csc_expand<<<grid_for(nnz), TPB, 0, st>>>(d_rowval, nnz, rpos, keyR, valE);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, need, keyR, posA, valE, a_csc, (int)nnz, 0, 32, st));
        CK(ensure_tmp(need));
        CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, keyR, posA, valE, a_csc, (int)nnz, 0, bits_for((uint64_t)m), st));
        k_inverse_perm<<<grid_for(nnz), TPB, 0, st>>>(a_csc, nnz, apos_of_csc);
        k_b_keys<<<grid_for(nnz), TPB, 0, st>>>(posA, a_csc, nnz, h->ft_row_of_pos, tile_incl, col_of, keyB);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, need, keyB, keyB_s, a_csc, b_csc, (int)nnz, 0, 64, st));
        CK(ensure_tmp(need));
        CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, keyB, keyB_s, a_csc, b_csc, (int)nnz, 0,
                                           32 + bits_for((uint64_t)n_tiles), st));
        k_b_tiles<<<grid_for(nnz), TPB, 0, st>>>(keyB_s, nnz, tileB);
    }
    pt.mark("fused: row sort + two entry sorts");
    uint32_t *segstart, *colstart, *segs_before, *cols_before, *run_pos;
    CK(sc.alloc(&segstart, nnz + 1)); CK(sc.alloc(&colstart, nnz + 1)); CK(sc.alloc(&segs_before, nnz + 1));
    CK(sc.alloc(&cols_before, nnz + 1)); CK(sc.alloc(&run_pos, nnz + 1));
    k_col_starts<<<grid_for(nnz + 1), TPB, 0, st>>>(tileB, b_csc, nnz, col_of, colstart);
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, colstart, cols_before, (int)(nnz + 1), st));
    CK(ensure_tmp(need));
    CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, colstart, cols_before, (int)(nnz + 1), st));
    if (nnz > 0) k_run_pos<<<grid_for(nnz), TPB, 0, st>>>(colstart, cols_before, nnz, run_pos);
    k_seg_starts<<<grid_for(nnz + 1), TPB, 0, st>>>(colstart, cols_before, run_pos, nnz, segstart);
    CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, segstart, segs_before, (int)(nnz + 1), st));

    // ---- per-tile headers, blob offsets
    FusedHdr *hdrs;
    uint64_t *blob_bytes, *blob_off;
    uint32_t *d_max;
    uint32_t *col0;
    CK(sc.alloc(&hdrs, n_tiles)); CK(sc.alloc(&blob_bytes, (size_t)n_tiles + 1)); CK(sc.alloc(&blob_off, (size_t)n_tiles + 1));
    CK(sc.alloc(&col0, n_tiles));
    CK(sc.alloc(&d_max, 8));
    CK(cudaMemsetAsync(d_max, 0, 32, st));
    CK(cudaMemsetAsync(blob_bytes, 0, sizeof(uint64_t) * ((size_t)n_tiles + 1), st));
    k_tile_meta<<<grid_for(n_tiles), TPB, 0, st>>>(n_tiles, tile_row0, row_ptr, len_sorted, segs_before, cols_before, hdrs, col0, blob_bytes, d_max);
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, blob_bytes, blob_off, (int)(n_tiles + 1), st));
    CK(ensure_tmp(need));
    CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, blob_bytes, blob_off, (int)(n_tiles + 1), st));
    uint32_t maxima[8];
    uint64_t total_bytes = 0;
    uint32_t n_parts = 0;
    CK(cudaMemcpyAsync(maxima, d_max, 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&total_bytes, blob_off + n_tiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&n_parts, segs_before + nnz, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    pt.mark("fused: runs + tile headers");
    // locality gate: the (tile, column) partials are written and read back once per step (32 B each at K = 8);
    // with rows in random order they would outweigh the matrix itself
    const char *force = getenv("POLEE_LAYOUT");
    const bool forced = force && std::string(force) == "fused";
    if (!forced && (double)n_parts * 64.0 > 0.35 * (double)total_bytes) return unsuitable();
    if (maxima[3] > 64u * 1024u) return unsuitable();  // a tile must fit a shared-memory stage
    if (maxima[1] > FT_MAX_C) return unsuitable();     // local column ids are 8 bits
    if (maxima[5] > FT_MAX_EA) return unsuitable();    // 16-bit slab offsets

    CK(polee::dmalloc((void **)&h->ft_blob, std::max<uint64_t>(total_bytes, 16)));
    CK(polee::dmalloc((void **)&h->ft_desc, sizeof(FusedTileDesc) * std::max<uint32_t>(n_tiles, 1)));
    CK(cudaMemsetAsync(h->ft_blob, 0, std::max<uint64_t>(total_bytes, 16), st));
    uint32_t *part_col, *part_col_sorted, *pid_iota, *col_cnt;
    CK(sc.alloc(&part_col, n_parts)); CK(sc.alloc(&part_col_sorted, n_parts)); CK(sc.alloc(&pid_iota, n_parts));
    CK(sc.alloc(&col_cnt, n));
    uint32_t *plist, *part_tile;
    CK(sc.alloc(&plist, n_parts)); CK(sc.alloc(&part_tile, n_parts));
    k_tile_desc<<<grid_for(n_tiles), TPB, 0, st>>>(n_tiles, hdrs, blob_off, blob_bytes, len_sorted, h->ft_desc, h->ft_blob);
    if (vals_ready_or_null) CK(cudaStreamWaitEvent(st, vals_ready_or_null, 0));
    if (nnz > 0) {
        k_pack_a<<<grid_for(nnz), TPB, 0, st>>>(nnz, posA, a_csc, h->ft_row_of_pos, tile_incl, row_ptr, hdrs, blob_off, d_nzval, h->ft_blob);
        k_pack_b<<<grid_for(nnz), TPB, 0, st>>>(nnz, tileB, b_csc, apos_of_csc, posA, col_of, segstart, colstart, segs_before,
                                                 cols_before, row_ptr, hdrs, col0, blob_off, d_nzval, h->ft_blob, part_col, part_tile);
    }
    pt.mark("fused: pack blobs");
    // ---- second stage work list: partial ids by column
    CK(cudaMemsetAsync(col_cnt, 0, sizeof(uint32_t) * n, st));
    if (n_parts > 0) {
        k_iota_u32<<<grid_for(n_parts), TPB, 0, st>>>(pid_iota, n_parts);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, need, part_col, part_col_sorted, pid_iota, plist, (int)n_parts, 0, 32, st));
        CK(ensure_tmp(need));
        CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, part_col, part_col_sorted, pid_iota, plist, (int)n_parts, 0,
                                           bits_for((uint64_t)n), st));
        k_count_keys<<<grid_for(n_parts), TPB, 0, st>>>(part_col, n_parts, col_cnt);
        k_pack_dest<<<grid_for(n_parts), TPB, 0, st>>>(n_parts, plist, part_tile, hdrs, blob_off, h->ft_blob);
    }
    std::vector<uint32_t> cnt(n);
    CK(cudaMemcpyAsync(cnt.data(), col_cnt, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
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
    CK(polee::dmalloc((void **)&h->ft_units, sizeof(FusedUnit) * std::max<size_t>(units.size(), 1)));
    CK(polee::dmalloc((void **)&h->ft_multi, sizeof(FusedMulti) * std::max<size_t>(multi.size(), 1)));
    CK(cudaMemcpyAsync(h->ft_units, units.data(), sizeof(FusedUnit) * units.size(), cudaMemcpyHostToDevice, st));
    if (!multi.empty())
        CK(cudaMemcpyAsync(h->ft_multi, multi.data(), sizeof(FusedMulti) * multi.size(), cudaMemcpyHostToDevice, st));
    if (d_ks) {
        CK(polee::dmalloc((void **)&h->ft_row_weight, sizeof(float) * m));
        k_ks_to_f32<<<grid_for(m), TPB, 0, st>>>(d_ks, h->ft_row_of_pos, m, h->ft_row_weight);
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    pt.mark("fused: second-stage list");
    h->ft_tiles = (int)n_tiles;
    h->ft_max_E = maxima[0]; h->ft_max_C = maxima[1]; h->ft_max_rows = maxima[2]; h->ft_max_blob = maxima[3];
    h->ft_blob_bytes = total_bytes;
    h->ft_parts = n_parts;
    h->ft_nunits = (int)units.size();
    h->ft_nmulti = (int)multi.size();
    h->ft_nlvl2 = (int)lvl2;
    h->fused = true;
    if (getenv("POLEE_SETUP_TIMING"))
        fprintf(stderr, "[polee setup] fused: %u tiles, %.1f MB blobs (%.2f B/entry), %u partials, %zu units, %zu multi, max E %u C %u rows %u blob %u segs %u EA %u\n",
                n_tiles, total_bytes / 1e6, nnz ? (double)total_bytes / nnz : 0.0, n_parts, units.size(), multi.size(),
                maxima[0], maxima[1], maxima[2], maxima[3], maxima[4], maxima[5]);
    return POLEE_OK;
}

// The general layouts ("split" / "fused") of m rows -- the whole matrix, or the rows the equivalence-class layout
// did not take.
static int setup_general_layouts(polee_handle *h, int64_t m, int64_t n, int64_t nnz, const uint32_t *d_colptr,
                                 const uint32_t *d_rowval, const float *d_nzval, const int64_t *d_ks,
                                 const std::vector<uint32_t> &colptr, cudaEvent_t vals_ready_or_null) {
    cudaStream_t st = h->stream;
    PhaseTimer pt(st);
    h->gm = m; h->gnnz = nnz;

    // Two layouts.  "split" (SELL slabs for K1 + re-sorted CSC for K2) streams the matrix twice and round-trips w
    // through HBM; both kernels are HBM-bound (0.84 / 0.82 of peak).  "fused" (row tiles, one pass, w stays in shared
    // memory) moves ~40 % of the bytes, builds in less than half the time and takes half the memory; it is bound by
    // the shared-memory pipe and the CTA barriers instead.  Measured at C3, likelihood pass of the rank-0 block of an
    // N-way partition, fused vs split: N = 1 0.750 vs 0.767 ms, 2: 0.397 vs 0.417, 4: 0.216 vs 0.245, 8: 0.127 vs 0.158.
    // Fused is therefore the default whenever the row order has the locality it needs (the builder declines
    // otherwise: position-sorted rows, as the reference produces them, have it); POLEE_LAYOUT=split|fused overrides;
    // exact_accumulation always uses the reference-order split kernels.
    {
        const char *lay = getenv("POLEE_LAYOUT");
        const bool want_fused = !h->o.exact_accumulation && !(lay && std::string(lay) == "split");
        if (want_fused) {
            int rc = setup_fused_from_device_csc(h, m, n, nnz, d_colptr, d_rowval, d_nzval, d_ks, colptr, vals_ready_or_null);
            if (rc) return rc;
            if (h->fused) {
                h->have_matrix = true;
                return POLEE_OK;
            }
        }
    }

    const int TPB = 256;
    auto grid_for = [&](int64_t work) { return (int)std::min<int64_t>((work + TPB - 1) / TPB, (int64_t)h->num_sms * 32); };

    Scratch sc;
    uint32_t *row_len, *keys_a, *keys_b, *vals_a, *vals_b;
    int *d_bad;
    CK(sc.alloc(&row_len, m));
    CK(sc.alloc(&d_bad, 1));
    CK(cudaMemsetAsync(row_len, 0, sizeof(uint32_t) * m, st));
    CK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    if (nnz > 0) k_count_rows<<<grid_for(nnz), TPB, 0, st>>>(d_rowval, nnz, row_len, m, d_bad);
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));

    // ---- row length classes
    uint32_t *d_lmax;
    CK(sc.alloc(&d_lmax, 1));
    size_t tmp_bytes = 0;
    void *d_tmp = nullptr;
    CK(cub::DeviceReduce::Max(nullptr, tmp_bytes, row_len, d_lmax, (int)m, st));
    CK(sc.alloc((char **)&d_tmp, tmp_bytes));
    CK(cub::DeviceReduce::Max(d_tmp, tmp_bytes, row_len, d_lmax, (int)m, st));
    uint32_t lmax = 0;
    CK(cudaMemcpyAsync(&lmax, d_lmax, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (bad) return h->fail(POLEE_EINVAL, "set_matrix: rowval out of range 1..m");
    pt.mark("row lengths + max");

    int *d_hist;
    CK(sc.alloc(&d_hist, (size_t)lmax + 1));
    {
        size_t hb = 0;
        void *ht = nullptr;
        CK(cub::DeviceHistogram::HistogramEven(nullptr, hb, row_len, d_hist, (int)lmax + 2, 0u, lmax + 1u, (int)m, st));
        CK(sc.alloc((char **)&ht, hb));
        CK(cub::DeviceHistogram::HistogramEven(ht, hb, row_len, d_hist, (int)lmax + 2, 0u, lmax + 1u, (int)m, st));
    }
    std::vector<int> hist(lmax + 1);
    CK(cudaMemcpyAsync(hist.data(), d_hist, sizeof(int) * (lmax + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));

    // classes in descending length; each padded to a multiple of 32 rows (one warp = one class)
    std::vector<uint32_t> cls_row_off_pad(lmax + 1, 0), cls_row_off_unpad(lmax + 1, 0), cls_stride(lmax + 1, 0);
    std::vector<uint64_t> cls_slab_off(lmax + 1, 0);
    std::vector<RowTile> tiles;
    uint64_t rows_pad = 0, rows_unpad = 0, slab = 0;
    for (int64_t L = lmax; L >= 0; --L) {
        uint32_t cnt = (uint32_t)hist[L];
        uint32_t stride = L > 0 ? (cnt + 31u) & ~31u : cnt;
        cls_row_off_pad[L] = (uint32_t)rows_pad;
        cls_row_off_unpad[L] = (uint32_t)rows_unpad;
        cls_stride[L] = stride;
        cls_slab_off[L] = slab;
        if (L > 0)
            for (uint32_t r0 = 0; r0 < cnt; r0 += ROW_TILE) {
                RowTile t;
                t.slab_off = slab + r0;
                t.stride = stride;
                t.len = (uint32_t)L;
                t.row0 = (uint32_t)rows_pad + r0;
                t.nrows = std::min<uint32_t>(ROW_TILE, cnt - r0);
                tiles.push_back(t);
            }
        rows_pad += stride;
        rows_unpad += cnt;
        slab += (uint64_t)L * stride;
    }
    if (rows_pad >= 0xFFFFFFFFull) return h->fail(POLEE_EINVAL, "set_matrix: too many rows");
    h->m_pad = (int64_t)rows_pad;
    h->sell_elems = (int64_t)slab;

    uint32_t *d_cls_pad, *d_cls_unpad, *d_cls_stride;
    uint64_t *d_cls_slab;
    CK(sc.alloc(&d_cls_pad, lmax + 1)); CK(sc.alloc(&d_cls_unpad, lmax + 1));
    CK(sc.alloc(&d_cls_stride, lmax + 1)); CK(sc.alloc(&d_cls_slab, lmax + 1));
    CK(cudaMemcpyAsync(d_cls_pad, cls_row_off_pad.data(), 4 * (lmax + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cls_unpad, cls_row_off_unpad.data(), 4 * (lmax + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cls_stride, cls_stride.data(), 4 * (lmax + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cls_slab, cls_slab_off.data(), 8 * (lmax + 1), cudaMemcpyHostToDevice, st));

    pt.mark("length classes (host)");
    // ---- stable sort of rows by descending length -> permutation
    const int64_t big = std::max<int64_t>(m, nnz);
    CK(sc.alloc(&keys_a, big)); CK(sc.alloc(&keys_b, big)); CK(sc.alloc(&vals_a, big)); CK(sc.alloc(&vals_b, big));
    k_make_len_keys<<<grid_for(m), TPB, 0, st>>>(row_len, m, lmax, keys_a, vals_a);
    size_t sort_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_a, keys_b, vals_a, vals_b, (int)big, 0, 32, st));
    void *d_sort = nullptr;
    CK(sc.alloc((char **)&d_sort, sort_bytes));
    CK(cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, keys_a, keys_b, vals_a, vals_b, (int)m, 0, bits_for(lmax), st));

    CK(polee::dmalloc((void **)&h->row_perm, sizeof(uint32_t) * m));
    uint32_t *len_perm, *row_start;
    CK(sc.alloc(&len_perm, rows_pad)); CK(sc.alloc(&row_start, rows_pad));
    CK(cudaMemsetAsync(len_perm, 0, sizeof(uint32_t) * std::max<uint64_t>(rows_pad, 1), st));
    k_assign_perm<<<grid_for(m), TPB, 0, st>>>(vals_b, row_len, m, d_cls_pad, d_cls_unpad, h->row_perm, len_perm);

    pt.mark("row sort + permutation");
    // ---- entries: sort by permuted row (stable from CSC order => ascending column inside a row)
    uint32_t *col_of;
    CK(sc.alloc(&col_of, nnz));
    uint32_t *d_colptr_own = nullptr;
    if (!d_colptr) {
        CK(sc.alloc(&d_colptr_own, n + 1));
        CK(cudaMemcpyAsync(d_colptr_own, colptr.data(), 4 * (n + 1), cudaMemcpyHostToDevice, st));
        d_colptr = d_colptr_own;
    }
    CK(polee::dmalloc((void **)&h->sell_idx, sizeof(uint32_t) * std::max<uint64_t>(slab, 1)));
    CK(polee::dmalloc((void **)&h->sell_val, sizeof(float) * std::max<uint64_t>(slab, 1)));
    CK(cudaMemsetAsync(h->sell_idx, 0, sizeof(uint32_t) * std::max<uint64_t>(slab, 1), st));
    CK(cudaMemsetAsync(h->sell_val, 0, sizeof(float) * std::max<uint64_t>(slab, 1), st));
    // +16 elements: K2's bulk copies round their byte count up to 16
    CK(polee::dmalloc((void **)&h->csc_row, sizeof(uint32_t) * (nnz + 16)));
    CK(polee::dmalloc((void **)&h->csc_val, sizeof(float) * (nnz + 16)));
    CK(cudaMemsetAsync(h->csc_row, 0, sizeof(uint32_t) * (nnz + 16), st));
    CK(cudaMemsetAsync(h->csc_val, 0, sizeof(float) * (nnz + 16), st));
    pt.mark("layout alloc + clear");
    if (nnz > 0) {
        k_entry_keys<<<grid_for(nnz), TPB, 0, st>>>(d_colptr, n, d_rowval, nnz, h->row_perm, col_of, keys_a, vals_a);
        CK(cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, keys_a, keys_b, vals_a, vals_b, (int)nnz, 0,
                                           bits_for(rows_pad), st));
        pt.mark("entry sort by row");
        // keys_b = sorted permuted rows, vals_b = entry ids; from here on nzval (and ks) are read
        if (vals_ready_or_null) CK(cudaStreamWaitEvent(st, vals_ready_or_null, 0));
        k_row_starts<<<grid_for(nnz), TPB, 0, st>>>(keys_b, nnz, row_start);
        // reuse keys_a / vals_a for the column sort
        k_fill_sell<<<grid_for(nnz), TPB, 0, st>>>(keys_b, vals_b, nnz, row_start, len_perm, d_cls_pad, d_cls_slab,
                                                    d_cls_stride, col_of, d_nzval, h->sell_idx, h->sell_val, keys_a,
                                                    vals_a);
        pt.mark("SELL fill");
        // stable sort by column of the (row', col)-ordered list -> (col, row') order
        uint32_t *keys_c, *vals_c;
        CK(sc.alloc(&keys_c, nnz)); CK(sc.alloc(&vals_c, nnz));
        CK(cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, keys_a, keys_c, vals_a, vals_c, (int)nnz, 0,
                                           bits_for((uint64_t)n), st));
        pt.mark("entry sort by column");
        k_fill_csc<<<grid_for(nnz), TPB, 0, st>>>(vals_c, nnz, keys_b, vals_b, d_nzval, h->csc_row, h->csc_val);
        pt.mark("CSC fill");
    }
    if (d_ks) {
        if (vals_ready_or_null) CK(cudaStreamWaitEvent(st, vals_ready_or_null, 0));
        CK(polee::dmalloc((void **)&h->row_weight, sizeof(float) * std::max<uint64_t>(rows_pad, 1)));
        CK(cudaMemsetAsync(h->row_weight, 0, sizeof(float) * std::max<uint64_t>(rows_pad, 1), st));
        k_row_weights<<<grid_for(m), TPB, 0, st>>>(d_ks, m, h->row_perm, h->row_weight);
    }

    // ---- K1 tiles, K2 segments
    h->n_row_tiles = (int)tiles.size();
    CK(polee::dmalloc((void **)&h->row_tiles, sizeof(RowTile) * std::max<size_t>(tiles.size(), 1)));
    if (!tiles.empty())
        CK(cudaMemcpyAsync(h->row_tiles, tiles.data(), sizeof(RowTile) * tiles.size(), cudaMemcpyHostToDevice, st));

    std::vector<ColSeg> segs;
    std::vector<MultiCol> multi;
    segs.reserve((size_t)(nnz / COL_SEG + n));
    for (int64_t j = 0; j < n; ++j) {
        uint32_t s = colptr[j] - 1, len = colptr[j + 1] - colptr[j];
        uint32_t ns = len == 0 ? 1 : (len + COL_SEG - 1) / COL_SEG;
        for (uint32_t q = 0; q < ns; ++q) {
            ColSeg sg;
            sg.start = s + q * COL_SEG;
            sg.len = std::min<uint32_t>(COL_SEG, len - q * COL_SEG);
            sg.col = (uint32_t)j;
            sg.slot = -1;
            segs.push_back(sg);
        }
    }
    // runs: consecutive segments of one column inside one CTA item (8 consecutive segments)
    constexpr size_t ITEM = 8;
    uint32_t slots = 0;
    for (size_t i = 0; i < segs.size();) {
        const uint32_t col = segs[i].col;
        size_t j = i;
        while (j < segs.size() && segs[j].col == col) ++j;      // [i, j) = all segments of this column
        size_t nruns = 0;
        const uint32_t first_slot = slots;
        for (size_t a = i; a < j;) {
            size_t b = std::min(j, (a / ITEM + 1) * ITEM);         // run = [a, b) inside one item
            segs[a].len |= (uint32_t)(b - a) << 16;
            ++nruns;
            a = b;
        }
        if (nruns > 1) {
            for (size_t a = i; a < j;) {
                size_t b = std::min(j, (a / ITEM + 1) * ITEM);
                segs[a].slot = (int32_t)slots++;
                a = b;
            }
            multi.push_back(MultiCol{col, first_slot, (uint32_t)nruns, 0});
        }
        i = j;
    }
    pt.mark("segments (host)");
    h->n_segs = (int)segs.size();
    h->n_multi = (int)multi.size();
    h->n_slots = (int)slots;
    CK(polee::dmalloc((void **)&h->segs, sizeof(ColSeg) * std::max<size_t>(segs.size(), 1)));
    CK(cudaMemcpyAsync(h->segs, segs.data(), sizeof(ColSeg) * segs.size(), cudaMemcpyHostToDevice, st));
    CK(polee::dmalloc((void **)&h->multi, sizeof(MultiCol) * std::max<size_t>(multi.size(), 1)));
    if (!multi.empty())
        CK(cudaMemcpyAsync(h->multi, multi.data(), sizeof(MultiCol) * multi.size(), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    pt.mark("descriptor upload");
    h->have_matrix = true;
    return POLEE_OK;
}

int setup_matrix_from_device_csc(polee_handle *h, int64_t m, int64_t n, const uint32_t *d_colptr,
                                 const uint32_t *d_rowval, const float *d_nzval, const int64_t *d_ks,
                                 const uint32_t *h_colptr_or_null, cudaEvent_t vals_ready_or_null) {
    release_matrix(h);
    if (m < 1 || n < 1) return h->fail(POLEE_EINVAL, "set_matrix: m and n must be >= 1");
    if (m > (int64_t)INT32_MAX) return h->fail(POLEE_EINVAL, "set_matrix: m exceeds 2^31 - 1 rows");
    cudaStream_t st = h->stream;
    PhaseTimer pt(st);
    pt.mark("inputs resident (H2D)");
    std::vector<uint32_t> colptr(n + 1);
    if (h_colptr_or_null)
        std::copy(h_colptr_or_null, h_colptr_or_null + n + 1, colptr.begin());
    else
        CK(polee::copy_sync(h->stream, colptr.data(), d_colptr, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost));
    if (colptr[0] != 1) return h->fail(POLEE_EINVAL, "set_matrix: colptr must be 1-based (colptr[1] == 1)");
    for (int64_t j = 0; j < n; ++j)
        if (colptr[j + 1] < colptr[j]) return h->fail(POLEE_EINVAL, "set_matrix: colptr is not non-decreasing");
    const int64_t nnz = (int64_t)colptr[n] - 1;
    if (nnz > (int64_t)INT32_MAX) return h->fail(POLEE_EINVAL, "set_matrix: more than 2^31 - 1 entries (partition the rows over ranks)");
    h->m = m; h->n = n; h->nnz = nnz;
    h->gm = 0; h->gnnz = 0;

    // Three layouts.  "ec" (equivalence classes, ec_setup.cu / ec_kernels.cu) is the default: rows with the same
    // transcript set are stored as dense blocks (values only) and both sparse products run as Float64 MMAs in one
    // pass.  Rows it does not take (small classes, very long rows) -- or all rows with POLEE_LAYOUT=split|fused, and
    // always with exact_accumulation (reference summation order, bit-identical frag_probs) -- use the general layouts:
    // "fused" (row tiles, one pass, Float32 inside a tile) where the row order has the locality it needs, else
    // "split" (SELL slabs for K1 + re-sorted CSC for K2, two passes, w through HBM).
    const char *lay = getenv("POLEE_LAYOUT");
    const bool want_ec = h->o.exact_accumulation != 1 && !(lay && (std::string(lay) == "split" || std::string(lay) == "fused"));
    uint32_t *d_colptr_own = nullptr;
    struct Own {
        uint32_t *&p;
        ~Own() { polee::dfree(p); }
    } own{d_colptr_own};
    if (!d_colptr) {
        CK(polee::dmalloc((void **)&d_colptr_own, sizeof(uint32_t) * (n + 1)));
        CK(cudaMemcpyAsync(d_colptr_own, colptr.data(), 4 * (n + 1), cudaMemcpyHostToDevice, st));
        d_colptr = d_colptr_own;
    }
    if (want_ec) {
        EcRest rest;
        int rc = setup_ec_from_device_csc(h, m, n, nnz, d_colptr, d_rowval, d_nzval, d_ks, vals_ready_or_null, &rest);
        if (rc) return rc;
        if (h->ec_tasks > 0) {
            if (rest.m == 0) {
                h->have_matrix = true;
                return POLEE_OK;
            }
            std::vector<uint32_t> rcolptr(n + 1);
            CK(polee::copy_sync(h->stream, rcolptr.data(), rest.colptr, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost));
            return setup_general_layouts(h, rest.m, n, rest.nnz, rest.colptr, rest.rowval, rest.nzval, rest.ks, rcolptr, nullptr);
        }
    }
    return setup_general_layouts(h, m, n, nnz, d_colptr, d_rowval, d_nzval, d_ks, colptr, vals_ready_or_null);
}

}  // namespace polee
