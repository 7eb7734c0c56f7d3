// Code components -> triplet dictionary on the GPU.  Replaces the host loops of
//   get_scanning_range_of_filtered_code_components   inference/_2_enumerate.jl:25-35
//   enumerate_triplets / insert_H!                    inference/_2_enumerate.jl:37-65
//   get_words / get_enriched_keys (the counting)      inference/_3_make_pfms.jl:3-26
// The reference walks every sequence's filtered code components sorted by position, forms all C(n,3) ordered triplets, keys
// them by (f1, f2, f3, d12, d13) and appends (sequence range index, position of the first component) to a dictionary entry per
// key; only keys with more than 10..200 entries are ever used.  20 000 sequences x 32 components are 10^8 triplets and 3*10^7
// distinct keys: minutes on the host, the longest stage of discover_motifs after the GPU took over training and scanning.
//
// Here: one warp per sequence range.  Pass 1 inserts every triplet's packed 64-bit key into an open-addressing hash table in HBM
// (64-bit CAS to claim a slot, 32-bit atomic add to count).  Only keys above the host's lowest count threshold are compacted;
// pass 2 re-enumerates and records, for those keys only, the index of their first insertion (the reference's Dictionary keeps
// insertion order, and get_enriched_keys' output order follows it); pass 3 re-enumerates once more and emits the dictionary
// values (range index, position) of the keys the host finally selected, tagged with their enumeration index so that the host
// restores insertion order with one sort of a few hundred thousand records.  Nothing is ever sorted on the device and the 10^8
// (key, value) pairs are never materialised.
#include "common.cuh"
#include <algorithm>

#define TR_MAXN 128                 // components per range (a range lies inside one sequence: <= q plus ties)
#define TR_WARPS 8

struct TrEntry { unsigned long long key; uint32_t count; uint32_t cand; };      // cand = 1 + index in the candidate list, 0 = none

struct mb200_triplets {
    int device = 0;
    int64_t n_codes = 0, n_ranges = 0, n_triplets = 0;
    std::vector<int32_t> h_start, h_stop;
    uint16_t *pos = nullptr, *fil = nullptr;          // device, sorted by position inside each range, 1-based values
    uint16_t *pos_in = nullptr, *fil_in = nullptr;
    int32_t *rs = nullptr, *re = nullptr;
    int64_t* base = nullptr;                          // enumeration index of each range's first triplet
    TrEntry* table = nullptr; uint64_t mask = 0; size_t table_bytes = 0;      // block from the ctx cache (mb_pool_alloc)
    mb200_key_count* cand = nullptr; int64_t n_cand = 0; uint32_t cand_min = 0; bool have_cand = false;
    uint32_t* sel = nullptr;
    unsigned long long* counters = nullptr;           // [0] generic counter, [1] error flag
};

__device__ __forceinline__ uint64_t tr_hash(uint64_t k) {          // murmur3 fmix64
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

// stable sort of every range by position (sort(by = x -> x[1]) on records already ordered by (fil, position): _2_enumerate.jl:52);
// ranks by counting, one warp per range
__global__ void __launch_bounds__(32 * TR_WARPS) trip_sort_kernel(const uint16_t* __restrict__ pin, const uint16_t* __restrict__ fin,
                                                                  const int32_t* __restrict__ rs, const int32_t* __restrict__ re, int64_t n_ranges,
                                                                  uint16_t* __restrict__ pout, uint16_t* __restrict__ fout) {
    __shared__ uint16_t sp[TR_WARPS][TR_MAXN];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * TR_WARPS + w;
    if (r >= n_ranges) return;
    const int a = rs[r], n = re[r] - a;
    for (int e = lane; e < n; e += 32) sp[w][e] = pin[a + e];
    __syncwarp();
    for (int e = lane; e < n; e += 32) {
        const uint16_t p = sp[w][e];
        int rank = 0;
        for (int o = 0; o < n; ++o) { const uint16_t q = sp[w][o]; rank += (q < p) || (q == p && o < e); }
        pout[a + rank] = (uint16_t)(p + 1);                  // 1-based like the reference's records
        fout[a + rank] = (uint16_t)(fin[a + e] + 1);
    }
}

// MODE 0: count keys.  MODE 1: first-insertion index of candidate keys.  MODE 2: emit values of selected keys.
template <int MODE>
__global__ void __launch_bounds__(32 * TR_WARPS) trip_enum_kernel(const uint16_t* __restrict__ pos, const uint16_t* __restrict__ fil,
                                                                  const int32_t* __restrict__ rs, const int32_t* __restrict__ re,
                                                                  const int64_t* __restrict__ base, int64_t n_ranges,
                                                                  TrEntry* __restrict__ table, uint64_t mask,
                                                                  mb200_key_count* __restrict__ cand, const uint32_t* __restrict__ sel,
                                                                  mb200_triplet_value* __restrict__ out, int64_t out_cap,
                                                                  unsigned long long* __restrict__ counters) {
    __shared__ uint16_t sp[TR_WARPS][TR_MAXN], sf[TR_WARPS][TR_MAXN];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * TR_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * TR_WARPS + w; r < n_ranges; r += nwarps) {
        const int a = rs[r], n = re[r] - a;
        if (n < 3) continue;
        __syncwarp();
        for (int e = lane; e < n; e += 32) { sp[w][e] = pos[a + e]; sf[w][e] = fil[a + e]; }
        __syncwarp();
        int64_t idx = base[r];                               // enumeration index of triplet (i, j, j+1)
        for (int i = 0; i < n - 2; ++i) {
            const uint64_t pi = sp[w][i], fi = sf[w][i];
            for (int j = i + 1; j < n - 1; ++j) {
                const uint64_t kij = fi | ((uint64_t)sf[w][j] << 8) | ((uint64_t)(sp[w][j] - pi) << 24);
                for (int k = j + 1 + lane; k < n; k += 32) {
                    const uint64_t key = kij | ((uint64_t)sf[w][k] << 16) | ((uint64_t)(sp[w][k] - pi) << 40);
                    uint64_t slot = tr_hash(key) & mask;
                    if (MODE == 0) {
                        while (true) {
                            const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&table[slot].key);
                            if (cur == key) break;
                            if (cur == 0ull) { const unsigned long long old = atomicCAS(&table[slot].key, 0ull, (unsigned long long)key); if (old == 0ull || old == key) break; }
                            slot = (slot + 1) & mask;
                        }
                        atomicAdd(&table[slot].count, 1u);
                    } else {
                        while (table[slot].key != key) slot = (slot + 1) & mask;          // present by construction
                        const uint32_t c = table[slot].cand;
                        if (c) {
                            const int64_t order = idx + (k - j - 1);
                            if (MODE == 1) atomicMin(reinterpret_cast<unsigned long long*>(&cand[c - 1].first), (unsigned long long)order);
                            else {
                                const uint32_t m = sel[c - 1];
                                if (m) {
                                    const unsigned long long o = atomicAdd(&counters[0], 1ull);
                                    if ((int64_t)o < out_cap) { mb200_triplet_value v; v.key_index = m - 1; v.range_index = (uint32_t)(r + 1); v.position = (uint32_t)pi; v.reserved = 0; v.order = (uint64_t)order; out[o] = v; }
                                }
                            }
                        }
                    }
                }
                idx += n - 1 - j;
            }
        }
    }
}

// keys with count > min_count: count them, then compact them into the candidate list and leave their list index in the table
__global__ void __launch_bounds__(256) trip_compact_kernel(TrEntry* __restrict__ table, uint64_t cap, uint32_t min_count,
                                                           mb200_key_count* __restrict__ cand, unsigned long long* __restrict__ counters) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap) return;
    TrEntry e = table[s];
    uint32_t c = 0;
    if (e.key != 0ull && e.count > min_count) {
        const unsigned long long i = atomicAdd(&counters[0], 1ull);
        if (cand) { mb200_key_count kc; kc.key = e.key; kc.count = e.count; kc.reserved = 0; kc.first = ~0ull; cand[i] = kc; c = (uint32_t)i + 1; }
    }
    if (cand) table[s].cand = c;
}

__global__ void __launch_bounds__(256) trip_mark_kernel(const TrEntry* __restrict__ table, uint64_t mask, const uint64_t* __restrict__ keys, int64_t n_keys,
                                                        uint32_t* __restrict__ sel, unsigned long long* __restrict__ counters) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_keys) return;
    const uint64_t key = keys[t];
    uint64_t slot = tr_hash(key) & mask;
    for (uint64_t probes = 0; probes <= mask; ++probes) {
        const unsigned long long cur = table[slot].key;
        if (cur == key) { const uint32_t c = table[slot].cand; if (c) sel[c - 1] = (uint32_t)t + 1; else atomicAdd(&counters[1], 1ull); return; }
        if (cur == 0ull) break;
        slot = (slot + 1) & mask;
    }
    atomicAdd(&counters[1], 1ull);                            // not a key of this dictionary (or below the candidate threshold)
}

// the hash table (2 x the triplets, GBs at 20 000 sequences) comes from the ctx's block cache: run_thru builds one dictionary per quantile
// (_g1_obtain_coutmats.jl:136-160), and cudaMalloc / cudaFree of a block that size cost ~30 ms each
static void trip_free(mb200_ctx* ctx, mb200_triplets* t) {
    if (!t) return;
    cudaFree(t->pos); cudaFree(t->fil); cudaFree(t->pos_in); cudaFree(t->fil_in); cudaFree(t->rs); cudaFree(t->re); cudaFree(t->base);
    if (t->table) mb_pool_free(ctx, t->table, t->table_bytes);
    cudaFree(t->cand); cudaFree(t->sel); cudaFree(t->counters);
    delete t;
}

extern "C" int32_t mb200_triplets_create(mb200_ctx* ctx, const uint16_t* position, const uint16_t* fil, const uint32_t* seq, int64_t n_codes,
                                         mb200_triplets** out, int64_t* n_ranges, int64_t* n_triplets) {
    if (!ctx) return MB200_E_INVALID;
    if (!out || n_codes < 0 || (n_codes > 0 && (!position || !fil || !seq))) MB_FAIL(ctx, MB200_E_INVALID, "triplets: bad arguments");
    *out = nullptr;
    MB_CUDA(ctx, cudaSetDevice(ctx->device));
    mb200_triplets* t = new mb200_triplets();
    t->device = ctx->device; t->n_codes = n_codes;
    // ranges, literally _2_enumerate.jl:25-35 with 1-based sequence ids: a range is closed whenever seq != cur_seq, cur_seq only counts
    // up by one per closed range, and the range of the last sequence is never closed
    {
        int64_t cur_seq = 1, start = 0;
        for (int64_t i = 0; i < n_codes; ++i) {
            if ((int64_t)seq[i] + 1 != cur_seq) { t->h_start.push_back((int32_t)start); t->h_stop.push_back((int32_t)i); start = i; ++cur_seq; }
        }
    }
    t->n_ranges = (int64_t)t->h_start.size();
    std::vector<int64_t> h_base(t->n_ranges + 1, 0);
    for (int64_t r = 0; r < t->n_ranges; ++r) {
        const int64_t n = t->h_stop[r] - t->h_start[r];
        if (n > TR_MAXN) { trip_free(ctx, t); MB_FAIL(ctx, MB200_E_UNSUPPORTED, "triplets: a range holds %lld code components (max %d)", (long long)n, TR_MAXN); }
        h_base[r + 1] = h_base[r] + (n >= 3 ? n * (n - 1) * (n - 2) / 6 : 0);
    }
    t->n_triplets = h_base[t->n_ranges];
    if (n_ranges) *n_ranges = t->n_ranges;
    if (n_triplets) *n_triplets = t->n_triplets;
    uint64_t cap = 1024;
    while (cap < 2 * (uint64_t)t->n_triplets && cap < (1ull << 31)) cap <<= 1;
    if ((uint64_t)t->n_triplets > cap / 2 + cap / 4) { trip_free(ctx, t); MB_FAIL(ctx, MB200_E_UNSUPPORTED, "triplets: %lld triplets exceed the hash table", (long long)t->n_triplets); }
    t->mask = cap - 1;
    const size_t nc = (size_t)std::max<int64_t>(n_codes, 1), nr = (size_t)std::max<int64_t>(t->n_ranges, 1);
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ok(cudaMalloc(&t->pos, nc * 2)); ok(cudaMalloc(&t->fil, nc * 2)); ok(cudaMalloc(&t->pos_in, nc * 2)); ok(cudaMalloc(&t->fil_in, nc * 2));
    ok(cudaMalloc(&t->rs, nr * 4)); ok(cudaMalloc(&t->re, nr * 4)); ok(cudaMalloc(&t->base, (nr + 1) * 8));
    t->table = reinterpret_cast<TrEntry*>(mb_pool_alloc(ctx, cap * sizeof(TrEntry), &t->table_bytes));
    if (!t->table) ok(cudaErrorMemoryAllocation);
    ok(cudaMalloc(&t->counters, 16));
    if (e != cudaSuccess) { trip_free(ctx, t); MB_FAIL(ctx, MB200_E_CUDA, "triplets: allocation failed: %s", cudaGetErrorString(e)); }
    cudaStream_t q = ctx->stream;
    ok(cudaMemsetAsync(t->table, 0, cap * sizeof(TrEntry), q));
    ok(cudaMemsetAsync(t->counters, 0, 16, q));
    if (n_codes) { ok(cudaMemcpyAsync(t->pos_in, position, (size_t)n_codes * 2, cudaMemcpyHostToDevice, q)); ok(cudaMemcpyAsync(t->fil_in, fil, (size_t)n_codes * 2, cudaMemcpyHostToDevice, q)); }
    if (t->n_ranges) {
        ok(cudaMemcpyAsync(t->rs, t->h_start.data(), (size_t)t->n_ranges * 4, cudaMemcpyHostToDevice, q));
        ok(cudaMemcpyAsync(t->re, t->h_stop.data(), (size_t)t->n_ranges * 4, cudaMemcpyHostToDevice, q));
        ok(cudaMemcpyAsync(t->base, h_base.data(), (size_t)(t->n_ranges + 1) * 8, cudaMemcpyHostToDevice, q));
        const unsigned blocks = (unsigned)((t->n_ranges + TR_WARPS - 1) / TR_WARPS);
        trip_sort_kernel<<<blocks, 32 * TR_WARPS, 0, q>>>(t->pos_in, t->fil_in, t->rs, t->re, t->n_ranges, t->pos, t->fil);
        const unsigned eb = (unsigned)std::min<int64_t>(blocks, (int64_t)ctx->sm_count * 8);
        trip_enum_kernel<0><<<eb, 32 * TR_WARPS, 0, q>>>(t->pos, t->fil, t->rs, t->re, t->base, t->n_ranges, t->table, t->mask, nullptr, nullptr, nullptr, 0, t->counters);
        ok(cudaGetLastError());
    }
    ok(cudaStreamSynchronize(q));
    if (e != cudaSuccess) { trip_free(ctx, t); MB_FAIL(ctx, MB200_E_CUDA, "triplets: %s", cudaGetErrorString(e)); }
    *out = t;
    return MB200_OK;
}

extern "C" int32_t mb200_triplets_destroy(mb200_ctx* ctx, mb200_triplets* t) {
    if (!t) return MB200_E_INVALID;
    if (ctx) cudaSetDevice(ctx->device);
    trip_free(ctx, t);
    return MB200_OK;
}

extern "C" int32_t mb200_triplets_ranges(mb200_ctx* ctx, const mb200_triplets* t, int32_t* start, int32_t* stop) {
    if (!ctx || !t || !start || !stop) return MB200_E_INVALID;
    std::copy(t->h_start.begin(), t->h_start.end(), start);
    std::copy(t->h_stop.begin(), t->h_stop.end(), stop);
    return MB200_OK;
}

extern "C" int32_t mb200_triplets_frequent(mb200_ctx* ctx, mb200_triplets* t, uint32_t min_count, mb200_key_count* out, int64_t cap, int64_t* n) {
    if (!ctx || !t || !n || cap < 0 || (cap > 0 && !out)) return MB200_E_INVALID;
    MB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t q = ctx->stream;
    if (!t->have_cand || t->cand_min != min_count) {
        cudaFree(t->cand); t->cand = nullptr; cudaFree(t->sel); t->sel = nullptr; t->n_cand = 0; t->have_cand = false;
        const uint64_t tcap = t->mask + 1;
        const unsigned blocks = (unsigned)((tcap + 255) / 256);
        unsigned long long h_cnt = 0;
        MB_CUDA(ctx, cudaMemsetAsync(t->counters, 0, 16, q));
        trip_compact_kernel<<<blocks, 256, 0, q>>>(t->table, tcap, min_count, nullptr, t->counters);          // count only
        MB_CUDA(ctx, cudaMemcpyAsync(&h_cnt, t->counters, 8, cudaMemcpyDeviceToHost, q));
        MB_CUDA(ctx, cudaStreamSynchronize(q));
        t->n_cand = (int64_t)h_cnt;
        MB_CUDA(ctx, cudaMalloc(&t->cand, (size_t)std::max<int64_t>(t->n_cand, 1) * sizeof(mb200_key_count)));
        MB_CUDA(ctx, cudaMalloc(&t->sel, (size_t)std::max<int64_t>(t->n_cand, 1) * 4));
        MB_CUDA(ctx, cudaMemsetAsync(t->sel, 0, (size_t)std::max<int64_t>(t->n_cand, 1) * 4, q));
        MB_CUDA(ctx, cudaMemsetAsync(t->counters, 0, 16, q));
        trip_compact_kernel<<<blocks, 256, 0, q>>>(t->table, tcap, min_count, t->cand, t->counters);
        if (t->n_cand && t->n_ranges) {
            const unsigned rb = (unsigned)((t->n_ranges + TR_WARPS - 1) / TR_WARPS);
            const unsigned eb = (unsigned)std::min<int64_t>(rb, (int64_t)ctx->sm_count * 8);
            trip_enum_kernel<1><<<eb, 32 * TR_WARPS, 0, q>>>(t->pos, t->fil, t->rs, t->re, t->base, t->n_ranges, t->table, t->mask, t->cand, nullptr, nullptr, 0, t->counters);
        }
        MB_CUDA(ctx, cudaGetLastError());
        MB_CUDA(ctx, cudaStreamSynchronize(q));
        t->have_cand = true; t->cand_min = min_count;
    }
    *n = t->n_cand;
    const int64_t m = std::min(cap, t->n_cand);
    if (m > 0) { MB_CUDA(ctx, cudaMemcpyAsync(out, t->cand, (size_t)m * sizeof(mb200_key_count), cudaMemcpyDeviceToHost, q)); MB_CUDA(ctx, cudaStreamSynchronize(q)); }
    return MB200_OK;
}

extern "C" int32_t mb200_triplets_values(mb200_ctx* ctx, mb200_triplets* t, const uint64_t* keys, int64_t n_keys, mb200_triplet_value* out, int64_t cap, int64_t* n) {
    if (!ctx || !t || !n || n_keys < 0 || cap < 0 || (n_keys > 0 && !keys) || (cap > 0 && !out)) return MB200_E_INVALID;
    if (!t->have_cand) MB_FAIL(ctx, MB200_E_INVALID, "triplets: call mb200_triplets_frequent first (values are kept for its keys only)");
    *n = 0;
    if (n_keys == 0 || t->n_cand == 0) { if (n_keys) MB_FAIL(ctx, MB200_E_INVALID, "triplets: no frequent keys in this dictionary"); return MB200_OK; }
    MB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t q = ctx->stream;
    uint64_t* d_keys = nullptr; mb200_triplet_value* d_out = nullptr;
    MB_CUDA(ctx, cudaMalloc(&d_keys, (size_t)n_keys * 8));
    cudaError_t e = cudaMalloc(&d_out, (size_t)std::max<int64_t>(cap, 1) * sizeof(mb200_triplet_value));
    if (e != cudaSuccess) { cudaFree(d_keys); MB_CUDA(ctx, e); }
    unsigned long long h_c[2] = {0, 0};
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ok(cudaMemcpyAsync(d_keys, keys, (size_t)n_keys * 8, cudaMemcpyHostToDevice, q));
    ok(cudaMemsetAsync(t->sel, 0, (size_t)t->n_cand * 4, q));
    ok(cudaMemsetAsync(t->counters, 0, 16, q));
    trip_mark_kernel<<<(unsigned)((n_keys + 255) / 256), 256, 0, q>>>(t->table, t->mask, d_keys, n_keys, t->sel, t->counters);
    const unsigned rb = (unsigned)((t->n_ranges + TR_WARPS - 1) / TR_WARPS);
    const unsigned eb = (unsigned)std::min<int64_t>(rb, (int64_t)ctx->sm_count * 8);
    trip_enum_kernel<2><<<eb, 32 * TR_WARPS, 0, q>>>(t->pos, t->fil, t->rs, t->re, t->base, t->n_ranges, t->table, t->mask, t->cand, t->sel, d_out, cap, t->counters);
    ok(cudaGetLastError());
    ok(cudaMemcpyAsync(h_c, t->counters, 16, cudaMemcpyDeviceToHost, q));
    ok(cudaStreamSynchronize(q));
    if (e == cudaSuccess && h_c[1] == 0) {
        *n = (int64_t)h_c[0];
        const int64_t m = std::min<int64_t>(cap, *n);
        if (m > 0) { ok(cudaMemcpyAsync(out, d_out, (size_t)m * sizeof(mb200_triplet_value), cudaMemcpyDeviceToHost, q)); ok(cudaStreamSynchronize(q)); }
    }
    cudaFree(d_keys); cudaFree(d_out);
    MB_CUDA(ctx, e);
    if (h_c[1]) MB_FAIL(ctx, MB200_E_INVALID, "triplets: %llu of the requested keys are not frequent keys of this dictionary", h_c[1]);
    return MB200_OK;
}
