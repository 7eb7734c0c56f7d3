// The unrolled CSC network as ONE persistent kernel (forward pass) for the reference's own step shape: a batch of 6 sequences.
//
// Reference: model.jl:330-395 (ADMM_XYZ, ADMM_DF, forward_pass_return_loss), restated in position space.
// The tape of csc.cu runs the same arithmetic as ~95 kernels of 2-25 us whose time is L2 round trips and launch gaps: six
// sequences cannot fill 148 SMs.  Here a sequence is owned by a CLUSTER of FZ_CL = 16 CTAs (CTA r owns a contiguous range of its c = Lb-7
// positions) and a whole pass is computed without leaving the kernel:
//   * per-sequence chains (recon -> corr_sig -> zy update, d build -> corr2d -> top-q -> tconv -> dual) need no barrier at all except
//     one cluster barrier for the per-sequence top-q: what a CTA needs from its neighbours (a halo of 7 or 11 positions) it
//     RECOMPUTES from tensors the neighbours published in global memory before the last barrier, instead of waiting for them;
//   * only the batch-wide statistics cross sequences: the median of the batch (two barriers over the group's 96 CTAs: merged
//     4096-bin histogram, then the median bin's candidates), the batch sums of the D and F gradients of ADMM_DF, the loss;
//   * the syntax filters F (115 KB), the dictionary D and the CTA's rows of z, y, alpha, beta, fx, theta stay in shared memory
//     across passes; every intermediate the reverse pass needs is also written to the tape's arena (fire-and-forget stores).
// Barriers over a group are counters in global memory (all CTAs are co-resident: cooperative launch); every sum that crosses
// CTAs is taken in a fixed order, so the step is deterministic.
#pragma once
#include "csc_kernels.cuh"
#include <cooperative_groups.h>

#define FZ_THREADS 512
#ifndef FZ_CL
#define FZ_CL 16                  // CTAs per sequence = cluster size.  16 is a non-portable cluster size (opt-in attribute, one cluster per GPC):
#endif                            // measured 0.65 -> 0.59 ms per step against 8 (make FZ_CL=8 builds that variant); handles that cannot place it keep the tape
#define FZ_M 50
#define FZ_M2 100
#define FZ_K 24
#define FZ_H 12
#define FZ_FL 8
#define FZ_FLEN 32
#define FZ_MAXPX 8
#define FZ_MAXPD 4
#define FZ_NHIST 4                // histogram levels a median call may use (window level, then bits 15..4, 3..0 or bits 30..19, 18..7, 6..0)
#ifndef FZ_KLO
#define FZ_KLO (103 << 7)         // window key = (float bits >> 16) - FZ_KLO: exponent field 103 (2^-24) .. 134 (2^8), 7 mantissa bits
#endif
#ifndef FZ_FORCE_GENERIC
#define FZ_FORCE_GENERIC 0        // test builds: 1 sends every median through the plain prefix levels
#endif
#ifndef FZ_CNTSEL
#define FZ_CNTSEL 512             // candidate sets up to this size are settled by counting (one candidate per thread), larger ones by radix passes
#endif
#define FZ_BINS 4096
#ifndef FZ_CALL
#define FZ_CALL                   // EXTRA_DEFS="-DFZ_CALL=__noinline__": the large phase routines as real calls (forward kernel 24 k -> 14 k SASS instructions).
#endif                            // Measured slower (0.548 -> 0.64 ms per step: Ctx / Smem then live in local memory, no per-site specialisation)
#ifndef FZ_CAND
#define FZ_CAND 2048
#endif

struct FzPass {                   // arena offsets (floats) of one ADMM_XYZ pass
    int64_t z_in, y_in, fx_in, al_in, be_in, rec, gz, gy, z_out, y_out, med, dd, g, x_in, x_out, fx_out, al_out, be_out, bits;
    int32_t xl_in, xl_out, i_eta, i_lam, i_rho, i_om;
};
struct FzDf {                     // one ADMM_DF pass; D_in/F_in have group stride 0 in pass 0 (the prepared filters)
    int64_t D_in, F_in, rec, Gm, Dn, e, Fg, Fn, nrm, thn;
    int32_t D_in_gs, F_in_gs, i_mu, i_kap, i_kaps, has_theta_out;
};
struct FzPlan {
    int32_t npx, npd, i_eta_w, i_lam_w, i_om_w, xl0, forward_only, pad0;
    int64_t sc, De, Fe, z0, y0, med0, zy0, g0, x0, fx0, bits0;
    FzPass px[FZ_MAXPX];
    int64_t zyF, medF, recL, fxL, loss;
    FzDf df[FZ_MAXPD];
};
struct FzBufs {
    float* data;                  // the tape's arena
    uint8_t* bits;
    int32_t* lcnt; uint16_t* lidx; float* lval;      // code lists [list][NS][LIST_CAP]
    const uint8_t* bases;         // [NS][Lb]
    unsigned int* bar;            // [G] barrier counters (zeroed before the launch)
    unsigned int* hist;           // [G][nmed][FZ_NHIST][FZ_BINS] merged histograms (zeroed before the launch)
    float* cand;                  // [G][nmed][FZ_CAND]
    unsigned int* cctl;           // [G][nmed][4]: candidates reserved, ~min bits above the bin
    float* part;                  // [G][B * FZ_CL][FZ_PART] per-CTA partial sums
};
#define FZ_PART 2048              // floats of partial-sum space per CTA (D gradient 1600, loss 2, ...)

// FZ_PROFILE builds (make FZ_PROFILE=1): thread 0 of CTA 0 accumulates clock64() per phase and prints the table at the end
#ifdef FZ_PROFILE
#define FZ_TDECL long long fz_t0 = clock64(), fz_acc[16] = {0}
#define FZ_T(id) do { if (threadIdx.x == 0) { const long long t = clock64(); fz_acc[id] += t - fz_t0; fz_t0 = t; } } while (0)
#else
#define FZ_TDECL
#define FZ_T(id)
#endif
#ifdef FZ_PROFILE
#define FZ_S0(c) do { if (threadIdx.x == 0) (c).subt = clock64(); } while (0)
#define FZ_S(c, id) do { if (threadIdx.x == 0) { const long long t = clock64(); (c).sub[id] += t - (c).subt; (c).subt = t; } } while (0)
#else
#define FZ_S0(c)
#define FZ_S(c, id)
#endif

namespace fz {
namespace cg = cooperative_groups;

struct Ctx {
    int n, r, g, gidx, ng, ncl;   // sequence, cluster rank, group, CTA index inside the group, CTAs per group, clusters per group
    int p0, p1, nr;               // own code rows [p0, p1)
    int i0, i1, ni;               // own x rows [i0, i1)
    int q1;                       // own base positions [p0, q1) for the signal (q1 = p1, the last CTA also takes the tail up to Lb)
    unsigned int epoch;           // barrier arrivals expected so far
    unsigned int* bar;
    long long tbar, tb1, tb2, tb3; int nbar; long long sub[24]; long long subt;     // FZ_PROFILE: clocks thread 0 spent inside group barriers (fence, arrive+poll, fence)
};

__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* p) {
    unsigned int v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
// global (L2) -> shared copy of n4 float4 with U loads per thread in flight: the destination is a generic pointer, so the compiler will not move
// a load above the previous iteration's store by itself, and a one-load-per-iteration loop pays one L2 round trip per 8 KB of the block
template <int U>
__device__ __forceinline__ void stage_f4(float4* dst, const float4* src, int n4) {
    for (int e0 = 0; e0 < n4; e0 += U * FZ_THREADS) {
        float4 v[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) { const int e = e0 + u * FZ_THREADS + threadIdx.x; if (e < n4) v[u] = __ldcg(src + e); }
        #pragma unroll
        for (int u = 0; u < U; ++u) { const int e = e0 + u * FZ_THREADS + threadIdx.x; if (e < n4) dst[e] = v[u]; }
    }
}

// barrier over the CTAs of one group; global writes before it are visible to every CTA of the group after it.  Hierarchical: the CTAs
// of a cluster meet at the hardware cluster barrier, only the cluster's rank-0 CTA arrives at / polls the counter in global memory (6
// arrivals on one address instead of 96: same-address atomics serialise in L2), a second cluster barrier releases the others.
// Ordering: the cluster barrier orders every CTA's writes before the leader's gpu-scope fence, which is cumulative over them; the leader's
// acquire fence after the poll and the second cluster barrier order them before every reader (which reads other CTAs' data with ld.cg).
#ifndef FZ_GB_ALLPOLL
#define FZ_GB_ALLPOLL 1           // 1: every CTA polls the counter itself after the leader's arrival; 0: only the leader polls and a second cluster barrier releases the others
#endif
__device__ __forceinline__ void group_barrier(Ctx& c) {
    cg::this_cluster().sync();
#if FZ_GB_ALLPOLL
    // One arrival per cluster (the leader's, after the cluster barrier above has ordered every CTA's writes before its gpu-scope fence), but every
    // CTA's thread 0 watches the counter: the release no longer costs a second cluster barrier (~500 clocks) after the leader's poll.  Reader
    // side: relaxed poll + acquire fence + block barrier in each CTA.
    if (threadIdx.x == 0) {
        c.epoch += (unsigned int)c.ncl;
        if (c.r == 0) { __threadfence(); asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(c.bar) : "memory"); }
        unsigned int spins = 0;
        while (ld_relaxed(c.bar) < c.epoch) { if (++spins > (1u << 26)) __trap(); }      // never hang the device
        __threadfence();
    }
    __syncthreads();
#else
    if (c.r == 0 && threadIdx.x == 0) {
#ifdef FZ_PROFILE
        const long long tb0 = clock64();
#endif
        c.epoch += (unsigned int)c.ncl;
        __threadfence();
#ifdef FZ_PROFILE
        const long long tb1 = clock64();
#endif
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(c.bar) : "memory");
        unsigned int spins = 0;
        while (ld_relaxed(c.bar) < c.epoch) { if (++spins > (1u << 26)) __trap(); }      // never hang the device
#ifdef FZ_PROFILE
        const long long tb2 = clock64();
#endif
        __threadfence();
#ifdef FZ_PROFILE
        const long long tb3 = clock64();
        c.tbar += tb3 - tb0; ++c.nbar; c.tb1 += tb1 - tb0; c.tb2 += tb2 - tb1; c.tb3 += tb3 - tb2;
#endif
    }
    cg::this_cluster().sync();
#endif
}
__device__ __forceinline__ void cluster_barrier() { cg::this_cluster().sync(); }

__device__ __forceinline__ float block_sum512(float v, float* s16) {      // fixed order; result in every thread
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s16[w] = v;
    __syncthreads();
    float r = 0.f;
    #pragma unroll
    for (int i = 0; i < FZ_THREADS / 32; ++i) r += s16[i];
    return r;
}
__device__ __forceinline__ int block_excl_scan512(int cnt, int* total, int* s_w /* 17 ints */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULLMASK, inc, o); if (lane >= o) inc += y; }
    __syncthreads();
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) { int run = 0; for (int i = 0; i < 16; ++i) { const int t = s_w[i]; s_w[i] = run; run += t; } s_w[16] = run; }
    __syncthreads();
    *total = s_w[16];
    return s_w[w] + inc - cnt;
}

// ---- shared-memory map (floats) --------------------------------------------------------------------------------------------
struct Smem {
    float *F, *D, *Dt, *Dr, *z, *y, *fx, *al, *be, *th, *zyF;   // persistent: filters (Dt/Dr: D as [tap][m][4] forward / reversed) and the CTA's rows of the state tensors
    float *A;                                           // work tile: d / zy rows [i0, i1 + 11) x 100, or z,y halo rows for recon
    float *sig;                                         // residual signal of base positions [p0, p1 + 7)
    float *w;                                           // work area: corr2d partials / top-q values / histograms / candidates
    float *gout;                                        // [ni][24] corr2d result of own rows
    int *li; float *lv; int *lc;                        // current code list of the sequence: flat index, value, count
    uint8_t* b;                                         // bases [p0 - 0, p1 + 7 + ...)
    float* red;                                         // 32 floats of reduction scratch
    int* iscr;                                          // 32 ints of scratch
};
__host__ __device__ inline int fz_rows(int c) { return (c + FZ_CL - 1) / FZ_CL; }
__host__ __device__ inline size_t fz_work_floats(int Lb) {
    const int c = Lb - FZ_FL + 1, l = c - FZ_H + 1;
    size_t a = (size_t)6144 + 64;                       // corr2d partials (<= 512 threads x 12)
    const size_t b = (size_t)l * FZ_K + 2048 + 64;      // top-q values + the 2048-bin histogram behind them
    const size_t h = (size_t)FZ_BINS + FZ_CAND;         // median histogram + candidates
    if (b > a) a = b;
    if (h > a) a = h;
    return a;
}
__host__ __device__ inline size_t fz_smem_bytes(int Lb) {
    const int c = Lb - FZ_FL + 1, R = fz_rows(c);
    size_t f = (size_t)FZ_H * FZ_M2 * FZ_K + 3 * FZ_FLEN * FZ_M;             // F, D, Dt, Dr
    f += (size_t)R * (FZ_M * 4 + FZ_M2 * 3);                                // z y al be | fx th zyF
    size_t A = (size_t)(R + FZ_H - 1) * FZ_M2, A2 = (size_t)2 * (R + 14) * FZ_M;
    f += (A > A2 ? A : A2);
    f += (size_t)4 * (R + 8);                                               // sig
    f += fz_work_floats(Lb);
    f += (size_t)R * FZ_K;                                                  // gout
    f += 3 * LIST_CAP + 8;                                                  // list
    f += 64;                                                                // red + iscr
    return f * 4 + (size_t)(R + 16) + 64;                                   // + bases
}

__device__ __forceinline__ void carve(Smem& s, float* base, int R, int Lb) {
    float* p = base;
    s.F = p; p += FZ_H * FZ_M2 * FZ_K;
    s.D = p; p += FZ_FLEN * FZ_M; s.Dt = p; p += FZ_FLEN * FZ_M; s.Dr = p; p += FZ_FLEN * FZ_M;
    s.z = p; p += R * FZ_M; s.y = p; p += R * FZ_M; s.al = p; p += R * FZ_M; s.be = p; p += R * FZ_M;
    s.fx = p; p += R * FZ_M2; s.th = p; p += R * FZ_M2; s.zyF = p; p += R * FZ_M2;
    const size_t A = (size_t)(R + FZ_H - 1) * FZ_M2, A2 = (size_t)2 * (R + 14) * FZ_M;
    s.A = p; p += (A > A2 ? A : A2);
    s.sig = p; p += 4 * (R + 8);
    s.w = p; p += fz_work_floats(Lb);
    s.gout = p; p += R * FZ_K;
    s.li = reinterpret_cast<int*>(p); p += LIST_CAP; s.lv = p; p += LIST_CAP; s.lc = reinterpret_cast<int*>(p); p += 8 + LIST_CAP;
    s.red = p; p += 32; s.iscr = reinterpret_cast<int*>(p); p += 32;
    s.b = reinterpret_cast<uint8_t*>(p);
}

// ---- batch median of the positive entries of (z, y) over the whole group; every CTA holds its rows in s.z / s.y -------------
// (create_ZY_mask, model.jl:194-204; Statistics.median: mean of the two middle values for an even count; no positives: -inf = no mask)
__device__ FZ_CALL float group_median(Ctx& c, const Smem& s, const FzBufs& B, int mi, int nmed) {
    unsigned int* lhist = reinterpret_cast<unsigned int*>(s.w);
    float* cand = s.w + FZ_BINS;
    unsigned int* wsum = reinterpret_cast<unsigned int*>(s.red);          // 32 uints
    unsigned int* res = reinterpret_cast<unsigned int*>(s.iscr);          // 4 uints (+ scratch)
    unsigned int* ghist0 = B.hist + ((size_t)c.g * nmed + mi) * FZ_NHIST * FZ_BINS;
    float* gcand = B.cand + ((size_t)c.g * nmed + mi) * FZ_CAND;
    unsigned int* gctl = B.cctl + ((size_t)c.g * nmed + mi) * 4;
    const int nv = c.nr * FZ_M;
    unsigned int prefix = 0, pmask = 0, krank = 0, npos = 0, cnt = 0;
    int shift = 16, nb = FZ_BINS;
    bool first = true, resolved = false, window = true;
    FZ_S0(c);
    // Level 0 bins the positives by WINDOW key: the top 16 bits (exponent + 7 mantissa bits) minus FZ_KLO, clamped to [0, 4095] -- 32 binades from
    // 2^-24 at a relative bin width of 2^-7, so the median's bin holds ~1 % of the entries and its candidates settle the rest in one exchange.
    // A median in an interior bin with few candidates (the usual case) is done after this level; an interior bin with many candidates (ties)
    // goes on with bits 15..4 and 3..0 under the bin's 16-bit prefix; an edge bin (values outside the window) restarts with the plain prefix
    // levels (bits 30..19, 18..7, 6..0) on the full rank.  Every decision is taken from the merged histogram: the same in every CTA.
    for (int lvl = 0; lvl < FZ_NHIST; ++lvl) {
        unsigned int* gh = ghist0 + (size_t)lvl * FZ_BINS;
        for (int i = threadIdx.x; i < nb / 4; i += FZ_THREADS) reinterpret_cast<uint4*>(lhist)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        for (int e = threadIdx.x; e < 2 * nv; e += FZ_THREADS) {
            const float f = e < nv ? s.z[e] : s.y[e - nv];
            const unsigned int b = __float_as_uint(f);
            if (window) { if (f > 0.f) atomicAdd(&lhist[min(max((int)(b >> 16) - FZ_KLO, 0), FZ_BINS - 1)], 1u); }
            else if (f > 0.f && (b & pmask) == prefix) atomicAdd(&lhist[(b >> shift) & (unsigned)(nb - 1)], 1u);
        }
        __syncthreads();
        FZ_S(c, 0);
        for (int i = threadIdx.x; i < nb / 4; i += FZ_THREADS) {
            const uint4 v = reinterpret_cast<const uint4*>(lhist)[i];
            if (v.x) atomicAdd(&gh[4 * i], v.x);
            if (v.y) atomicAdd(&gh[4 * i + 1], v.y);
            if (v.z) atomicAdd(&gh[4 * i + 2], v.z);
            if (v.w) atomicAdd(&gh[4 * i + 3], v.w);
        }
        FZ_S(c, 1);
        group_barrier(c);
        FZ_S(c, 2);
        for (int i = threadIdx.x; i < nb / 4; i += FZ_THREADS) reinterpret_cast<uint4*>(lhist)[i] = __ldcg(reinterpret_cast<const uint4*>(gh) + i);
        __syncthreads();
        FZ_S(c, 3);
        block_find_bin(lhist, nb, krank, first, wsum, res);
        if (first) {
            npos = res[3];
            if (npos == 0) break;
            krank = (npos & 1u) ? npos / 2 : npos / 2 - 1;
            first = false;
        }
        const unsigned int bin = res[0], below = res[1], inbin = res[2];
        __syncthreads();
        if (window) {
            window = false;
            if (bin >= 1u && bin <= (unsigned)(FZ_BINS - 2) && !FZ_FORCE_GENERIC) {
                krank -= below; cnt = inbin;
                prefix = (bin + (unsigned)FZ_KLO) << 16; pmask = 0xffff0000u; shift = 16;
                if (cnt <= FZ_CAND) break;
                shift = 4; nb = FZ_BINS;
            } else { prefix = 0; pmask = 0; shift = 19; nb = FZ_BINS; cnt = npos; }      // krank stays the rank among all positives
            continue;
        }
        krank -= below; cnt = inbin;
        prefix |= bin << shift; pmask |= (unsigned)(nb - 1) << shift;
        if (cnt <= FZ_CAND) break;
        if (shift == 0) { resolved = true; break; }
        if (shift == 19) { shift = 7; nb = FZ_BINS; } else if (shift == 7) { shift = 0; nb = 128; } else { shift = 0; nb = 16; }
    }
    FZ_S(c, 4);
#ifdef FZ_PROFILE_MEDIAN
    if (blockIdx.x == 0 && threadIdx.x == 0) printf("[fzm] median %d: %u positives, %u candidates, %d low bits open, window bin %d\n", mi, npos, cnt, shift, (int)(prefix >> 16) - FZ_KLO);
#endif
    float med = -INFINITY;
    if (npos > 0) {                                                        // group-uniform
        // publish this CTA's entries of the median's bin, and the smallest entry above the bin
        unsigned int* lc = lhist;                                          // local candidate bits
        if (threadIdx.x == 0) { res[0] = 0; res[1] = 0; }
        __syncthreads();
        unsigned int mxinv = 0;                                            // max of ~bits = min of bits above the bin (0: none)
        for (int e0 = 0; e0 < 2 * nv; e0 += FZ_THREADS) {
            const int e = e0 + threadIdx.x;
            const float f = e < 2 * nv ? (e < nv ? s.z[e] : s.y[e - nv]) : 0.f;
            const unsigned int b = __float_as_uint(f);
            if (f > 0.f && (b & pmask) > prefix) mxinv = max(mxinv, ~b);
            if (!resolved && f > 0.f && (b & pmask) == prefix) lc[atomicAdd(&res[0], 1u)] = b;
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) mxinv = max(mxinv, __shfl_xor_sync(FULLMASK, mxinv, o));
        if ((threadIdx.x & 31) == 0 && mxinv) atomicMax(&res[1], mxinv);
        __syncthreads();
        const unsigned int nc = res[0];
        if (threadIdx.x == 0) {
            res[2] = nc ? atomicAdd(&gctl[0], nc) : 0u;
            if (res[1]) atomicMax(&gctl[1], res[1]);
        }
        __syncthreads();
        const unsigned int base = res[2];
        for (unsigned int i = threadIdx.x; i < nc; i += FZ_THREADS) gcand[base + i] = __uint_as_float(lc[i]);
        FZ_S(c, 5);
        group_barrier(c);
        const unsigned int above_inv = __ldcg(&gctl[1]);
        FZ_S(c, 6);
        if (!resolved) for (unsigned int i = threadIdx.x; i < cnt; i += FZ_THREADS) cand[i] = __ldcg(&gcand[i]);
        __syncthreads();
        const unsigned int kin = krank;
        unsigned int v1b = prefix;
        // few candidates (the usual case with the window level: ~100): every thread ranks ONE candidate against all others (broadcast 16-byte
        // reads); the thread that holds rank krank also knows how many entries are <= it and the next larger one
        const bool cntsel = !resolved && cnt <= FZ_CNTSEL;
        unsigned int cs_le = 0, cs_next = 0x7f800000u;
        if (cntsel) {
            for (unsigned int i = cnt + threadIdx.x; i < ((cnt + 3u) & ~3u); i += FZ_THREADS) cand[i] = __uint_as_float(0x7f800000u);      // pad: +inf ranks above everything
            __syncthreads();
            if (threadIdx.x < cnt) {
                const unsigned int mine = __float_as_uint(cand[threadIdx.x]);
                unsigned int less = 0, eq = 0, nxt = 0x7f800000u;
                for (unsigned int j = 0; j < cnt; j += 4) {
                    const float4 o4 = *reinterpret_cast<const float4*>(cand + j);
                    const unsigned int o[4] = {__float_as_uint(o4.x), __float_as_uint(o4.y), __float_as_uint(o4.z), __float_as_uint(o4.w)};
                    #pragma unroll
                    for (int u = 0; u < 4; ++u) { less += o[u] < mine; eq += o[u] == mine; if (o[u] > mine) nxt = min(nxt, o[u]); }
                }
                if (less <= krank && krank < less + eq) { res[4] = mine; res[5] = less + eq; res[6] = nxt; }      // equal candidates write equal values
            }
            __syncthreads();
            v1b = res[4]; cs_le = res[5]; cs_next = res[6];
        } else if (!resolved) {
            unsigned int lowfix = 0, lowmask = 0;
            int rem = shift;
            while (rem > 0) {
                const int nbits = rem < 10 ? rem : 10, sh = rem - nbits, nbin = 1 << nbits;
                for (int i = threadIdx.x; i < nbin; i += FZ_THREADS) lhist[i] = 0;
                __syncthreads();
                for (unsigned e0 = 0; e0 < cnt; e0 += FZ_THREADS) {
                    const unsigned e = e0 + threadIdx.x;
                    const unsigned int b = e < cnt ? __float_as_uint(cand[e]) : 0u;
                    hist_add(lhist, e < cnt && (b & lowmask) == lowfix, (b >> sh) & (unsigned)(nbin - 1));
                }
                __syncthreads();
                block_find_bin(lhist, nbin, krank, false, wsum, res);
                lowfix |= res[0] << sh; lowmask |= (unsigned)(nbin - 1) << sh; krank -= res[1];
                __syncthreads();
                rem = sh;
            }
            v1b |= lowfix;
        }
        FZ_S(c, 7);
        const float v1 = __uint_as_float(v1b);
        if (npos & 1u) med = v1;
        else {
            // second middle value: a copy of v1, else the next candidate, else the smallest entry above the bin
            unsigned int le = 0, mn2 = 0x7f800000u;
            if (cntsel) { if (threadIdx.x == 0) { le = cs_le; mn2 = cs_next; } }
            else if (!resolved)
                for (unsigned e = threadIdx.x; e < cnt; e += FZ_THREADS) { const unsigned int b = __float_as_uint(cand[e]); if (b <= v1b) ++le; else mn2 = min(mn2, b); }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) { le += __shfl_xor_sync(FULLMASK, le, o); mn2 = min(mn2, __shfl_xor_sync(FULLMASK, mn2, o)); }
            __syncthreads();
            if (threadIdx.x == 0) { res[0] = 0; res[1] = 0x7f800000u; }
            __syncthreads();
            if ((threadIdx.x & 31) == 0) { if (le) atomicAdd(&res[0], le); atomicMin(&res[1], mn2); }
            __syncthreads();
            const unsigned int le_all = resolved ? cnt : res[0];
            const unsigned int v2b = (le_all >= kin + 2) ? v1b : (res[1] != 0x7f800000u ? res[1] : ~above_inv);
            med = v1 * 0.5f + __uint_as_float(v2b) * 0.5f;               // Statistics.middle(a, b) = a/2 + b/2
            __syncthreads();
        }
    }
    FZ_S(c, 8);
    return med;
}

// ---- D layer ---------------------------------------------------------------------------------------------------------------
// stage rows [lo, hi) of the sequence's z, y (global, [c][M]) into s.A as zh | yh (row-major [hi-lo][50] each)
__device__ __forceinline__ void stage_zy_halo(const Smem& s, const float* zg, const float* yg, int lo, int hi) {
    const int n = (hi - lo) * FZ_M;
    float* zh = s.A; float* yh = s.A + n;
    for (int e = threadIdx.x; e < n; e += FZ_THREADS) { zh[e] = __ldcg(zg + (size_t)lo * FZ_M + e); yh[e] = __ldcg(yg + (size_t)lo * FZ_M + e); }
}
// Dt[j][m][a] = D[4j+a][m], Dr[j][m][a] = D[31-4j-a][m]: the four nucleotide taps of a filter position as one 16-byte load
__device__ __forceinline__ void build_Dt(const Smem& s) {
    for (int o = threadIdx.x; o < FZ_FL * FZ_M; o += FZ_THREADS) {
        const int j = o / FZ_M, m = o - j * FZ_M;
        float4 f, r;
        f.x = s.D[(4 * j + 0) * FZ_M + m]; f.y = s.D[(4 * j + 1) * FZ_M + m]; f.z = s.D[(4 * j + 2) * FZ_M + m]; f.w = s.D[(4 * j + 3) * FZ_M + m];
        r.x = s.D[(31 - 4 * j) * FZ_M + m]; r.y = s.D[(30 - 4 * j) * FZ_M + m]; r.z = s.D[(29 - 4 * j) * FZ_M + m]; r.w = s.D[(28 - 4 * j) * FZ_M + m];
        reinterpret_cast<float4*>(s.Dt)[o] = f; reinterpret_cast<float4*>(s.Dr)[o] = r;
    }
}
// recon of base positions [q0, q1): rec[4q+a] = sum_{j<8} sum_m z[q-j][m] D[4j+a][m] + y[q-j][m] D[31-4j-a][m]  (model.jl:238-239,276-277,313-314)
// from the staged rows [lo, hi); half a warp per base position, lanes over m, four taps per 16-byte filter load.  Writes
// s.sig[t - 4 q0] = rec + sgn*S and, for own positions [wq0, wq1), rec to global (may be null)
__device__ FZ_CALL void recon_rows(const Ctx& c, const Smem& s, int lo, int hi, int q0, int q1, float sgn, float* rec_g, int wq0, int wq1, int Lb) {
    const int sub = threadIdx.x & 15, hw = threadIdx.x >> 4;
    const float* zh = s.A; const float* yh = s.A + (hi - lo) * FZ_M;
    const int cc = Lb - FZ_FL + 1;
    const float4* Dt4 = reinterpret_cast<const float4*>(s.Dt); const float4* Dr4 = reinterpret_cast<const float4*>(s.Dr);
    for (int qb = q0; qb < q1; qb += FZ_THREADS / 16) {             // every lane runs the same number of trips (full-warp shuffles below)
        const int q = qb + hw;
        const bool live = q < q1;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        #pragma unroll
        for (int j = 0; j < FZ_FL; ++j) {
            const int p = q - j;
            if (live && p >= 0 && p < cc) {
                #pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m = sub + 16 * i;
                    if (m < FZ_M) {
                        const float zv = zh[(p - lo) * FZ_M + m], yv = yh[(p - lo) * FZ_M + m];
                        const float4 f = Dt4[j * FZ_M + m], r = Dr4[j * FZ_M + m];
                        a0 += zv * f.x + yv * r.x; a1 += zv * f.y + yv * r.y; a2 += zv * f.z + yv * r.z; a3 += zv * f.w + yv * r.w;
                    }
                }
            }
        }
        #pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(FULLMASK, a0, o); a1 += __shfl_xor_sync(FULLMASK, a1, o);
            a2 += __shfl_xor_sync(FULLMASK, a2, o); a3 += __shfl_xor_sync(FULLMASK, a3, o);
        }
        if (live && sub == 0) {
            if (rec_g && q >= wq0 && q < wq1) *reinterpret_cast<float4*>(rec_g + 4 * q) = make_float4(a0, a1, a2, a3);
            float v[4] = {a0, a1, a2, a3};
            if (sgn != 0.f) v[s.b[q - c.p0]] += sgn;
            *reinterpret_cast<float4*>(s.sig + 4 * (q - q0)) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// ---- F layer ---------------------------------------------------------------------------------------------------------------
// out[i][k] = sum_{a<12} sum_{j<100} A[i+a][j] F[a][j][k] for the CTA's ni rows; A rows [i0, i1+11) in s.A; result in s.gout
__device__ FZ_CALL void corr2d_rows(const Smem& s, const float* Fm, int ni) {
    const int ntr = (ni + 2) / 3, ntile = ntr * 6;
    const int nslice = ntile ? min(FZ_THREADS / ntile, 64) : 0;
    const int tile = ntile ? threadIdx.x % ntile : 0, slice = ntile ? threadIdx.x / ntile : 0;
    const int tr = tile / 6, kg = tile - tr * 6;
    const int E = FZ_H * FZ_M2;
    if (ntile && slice < nslice) {
        float acc[3][4];
        #pragma unroll
        for (int r = 0; r < 3; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
        // slices are cut at multiples of 4 in e = a*100 + j: rows i..i+h-1 of A are one contiguous run in e, so four consecutive taps of a
        // row are ONE 16-byte load (shared memory bandwidth, not the FMA pipe, bounds this loop)
        const int e_lo = 4 * (int)((long long)(E / 4) * slice / nslice), e_hi = 4 * (int)((long long)(E / 4) * (slice + 1) / nslice);
        const float* Ab = s.A + (3 * tr) * FZ_M2;                 // row (3tr + rr + a), column j  ->  Ab[rr*100 + e]
        #pragma unroll 2
        for (int e = e_lo; e < e_hi; e += 4) {
            float4 av[3];
            #pragma unroll
            for (int r = 0; r < 3; ++r) av[r] = *reinterpret_cast<const float4*>(Ab + r * FZ_M2 + e);
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 f = *reinterpret_cast<const float4*>(Fm + (size_t)(e + u) * FZ_K + 4 * kg);
                #pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float a1 = u == 0 ? av[r].x : u == 1 ? av[r].y : u == 2 ? av[r].z : av[r].w;
                    acc[r][0] += a1 * f.x; acc[r][1] += a1 * f.y; acc[r][2] += a1 * f.z; acc[r][3] += a1 * f.w;
                }
            }
        }
        float* dst = s.w + ((size_t)slice * (3 * ntr) + 3 * tr) * FZ_K + 4 * kg;
        #pragma unroll
        for (int r = 0; r < 3; ++r) *reinterpret_cast<float4*>(dst + r * FZ_K) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < ni * FZ_K; o += FZ_THREADS) {
        float v = 0.f;
        for (int sl = 0; sl < nslice; ++sl) v += s.w[(size_t)sl * (3 * ntr) * FZ_K + o];
        s.gout[o] = v;
    }
    __syncthreads();
}

// out[i][j] = sum over a code list: v * F[i - iq][j][kq], 0 <= i - iq < 12, for own rows [p0, p1)   (model.jl:229,263,294,316,370).
// cnt <= LIST_CAP: a warp first builds, per own row, the ordered list of the codes that reach it (offset of F[a][.][kq], value) in s.w;
// then one thread per output (row, j) walks only those.  cnt <= lcap (a longer list held in li/lv): every output walks the whole list.
// Otherwise (forward only: more codes than the list holds) the dense tensor xg is read from global.  out: [nr][100].  Uses s.w.
// Filter element (a, j, k) sits at Fm[a*sa + j*sj + k*sk]: (2M*K, K, 1) for the [a][j][k] layout of F, (2M, 1, h*2M) for a k-major copy.
__device__ FZ_CALL void tconv_list(const Ctx& c, const Smem& s, const float* Fm, const int* li, const float* lv, int cnt, int lcap, const float* xg, int l, float* out,
                           int sa = FZ_M2 * FZ_K, int sj = FZ_K, int sk = 1) {
    const int R = c.nr;
    int* rcnt = reinterpret_cast<int*>(s.w); int* roff = rcnt + 32; float* rval = s.w + 32 + 32 * LIST_CAP;      // rows <= 32
    if (cnt <= LIST_CAP) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int il = warp; il < R; il += FZ_THREADS / 32) {
            const int i = c.p0 + il;
            int run = 0;
            for (int q0 = 0; q0 < cnt; q0 += 32) {
                const int q = q0 + lane;
                int a = -1, kq = 0; float v = 0.f;
                if (q < cnt) { const int e = li[q], iq = e / FZ_K; kq = e - iq * FZ_K; a = i - iq; v = lv[q]; }
                const bool hit = a >= 0 && a < FZ_H;
                const unsigned mk = __ballot_sync(FULLMASK, hit);
                if (hit) { const int o = run + __popc(mk & ((1u << lane) - 1u)); roff[il * LIST_CAP + o] = a * sa + kq * sk; rval[il * LIST_CAP + o] = v; }
                run += __popc(mk);
            }
            if (lane == 0) rcnt[il] = run;
        }
        __syncthreads();
        for (int o = threadIdx.x; o < R * FZ_M2; o += FZ_THREADS) {
            const int il = o / FZ_M2, j = o - il * FZ_M2;
            const int n = rcnt[il];
            float acc = 0.f;
            for (int q = 0; q < n; ++q) acc += rval[il * LIST_CAP + q] * Fm[roff[il * LIST_CAP + q] + j * sj];
            out[o] = acc;
        }
    } else if (cnt <= lcap) {
        for (int o = threadIdx.x; o < R * FZ_M2; o += FZ_THREADS) {
            const int il = o / FZ_M2, j = o - il * FZ_M2, i = c.p0 + il;
            float acc = 0.f;
            for (int q = 0; q < cnt; ++q) {
                const int e = li[q], iq = e / FZ_K, kq = e - iq * FZ_K, a = i - iq;
                if (a >= 0 && a < FZ_H) acc += lv[q] * Fm[a * sa + j * sj + kq * sk];
            }
            out[o] = acc;
        }
    } else {
        for (int o = threadIdx.x; o < R * FZ_M2; o += FZ_THREADS) {
            const int il = o / FZ_M2, j = o - il * FZ_M2, i = c.p0 + il;
            float acc = 0.f;
            const int a_lo = max(0, i - l + 1), a_hi = min(FZ_H - 1, i);
            for (int a = a_lo; a <= a_hi; ++a)
                for (int k = 0; k < FZ_K; ++k) { const float xv = __ldcg(xg + (size_t)(i - a) * FZ_K + k); if (xv != 0.f) acc += xv * Fm[a * sa + j * sj + k * sk]; }
            out[o] = acc;
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void tconv_rows(const Ctx& c, const Smem& s, const float* Fm, const float* xg, int l, float* out) {
    tconv_list(c, s, Fm, s.li, s.lv, s.lc[0], LIST_CAP, xg, l, out);
}

// per-sequence top-q (generate_bitmat / project_X, model.jl:181-192) on v = xprev + om * g over all l*K entries, redundantly in every CTA of
// the cluster; writes own rows of x and the bitmap, rank 0 writes the ordered code list; leaves the list in s.li / s.lv / s.lc
__device__ FZ_CALL void topq_all(Ctx& c, const Smem& s, const float* g_g, const float* xprev_g, bool have_prev, float om, int l, int q,
                         float* xout_g, uint8_t* bits_g, int32_t* lcnt_g, uint16_t* lidx_g, float* lval_g) {
    const int E = l * FZ_K;
    float* sv = s.w;
    FZ_S0(c);
    unsigned int* hist = reinterpret_cast<unsigned int*>(s.w + E);        // fz_work_floats() reserves 256 bins behind the values
    // values: previous codes come from the list when it is complete (cheaper than re-reading the dense tensor)
    const int pcnt = have_prev ? s.lc[0] : 0;
    if (have_prev && pcnt > LIST_CAP) { for (int e = threadIdx.x; e < E; e += FZ_THREADS) sv[e] = __ldcg(xprev_g + e) + om * __ldcg(g_g + e); }
    else {
        for (int e0 = 0; e0 < E; e0 += 4 * FZ_THREADS) {                 // four loads per thread in flight (see stage_f4)
            float gv[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) { const int e = e0 + u * FZ_THREADS + threadIdx.x; gv[u] = e < E ? __ldcg(g_g + e) : 0.f; }
            #pragma unroll
            for (int u = 0; u < 4; ++u) { const int e = e0 + u * FZ_THREADS + threadIdx.x; if (e < E) sv[e] = 0.f + om * gv[u]; }
        }
        __syncthreads();
        if (threadIdx.x < pcnt) { const int e = s.li[threadIdx.x]; sv[e] = s.lv[threadIdx.x] + om * __ldcg(g_g + e); }
    }
    FZ_S(c, 10);
    unsigned int* ctl = reinterpret_cast<unsigned int*>(s.iscr);          // [0] prefix / key of the q-th largest, [1] rank, [2] candidates
    unsigned int* wsum = reinterpret_cast<unsigned int*>(s.red);
    unsigned int* res = ctl + 4;
    // one pass over a 2048-bin histogram of the top 11 key bits (sign, exponent, 2 mantissa bits): the bin of the q-th largest value holds a
    // handful of entries, whose exact order is then settled by counting.  Many entries in that bin (ties): the byte-wise radix select below.
    for (int i = threadIdx.x; i < 2048; i += FZ_THREADS) hist[i] = 0;
    if (threadIdx.x == 0) { ctl[0] = 0; ctl[1] = (unsigned int)(E - q); ctl[2] = 0; }
    __syncthreads();
    for (int e0 = 0; e0 < E; e0 += FZ_THREADS) {
        const int e = e0 + threadIdx.x;
        const unsigned int kk = e < E ? fkey(sv[e]) : 0u;
        hist_add(hist, e < E, kk >> 21);
    }
    __syncthreads();
    FZ_S(c, 11);
    block_find_bin(hist, 2048, (unsigned int)(E - q), false, wsum, res);
    const unsigned int tbin = res[0], kr = (unsigned int)(E - q) - res[1], cntb = res[2];
    __syncthreads();
    FZ_S(c, 12);
    if (cntb <= FZ_THREADS) {
        unsigned int* cd = hist;                                            // the histogram is no longer needed
        for (int e0 = 0; e0 < E; e0 += FZ_THREADS) {
            const int e = e0 + threadIdx.x;
            if (e < E) { const unsigned int kk = fkey(sv[e]); if ((kk >> 21) == tbin) cd[atomicAdd(&ctl[2], 1u)] = kk; }
        }
        __syncthreads();
        if (threadIdx.x < cntb) {
            const unsigned int mine = cd[threadIdx.x];
            unsigned int less = 0, eq = 0;
            for (unsigned int i = 0; i < cntb; ++i) { const unsigned int o = cd[i]; less += o < mine; eq += o == mine; }
            if (less <= kr && kr < less + eq) ctl[0] = mine;              // every thread holding that value writes the same bits
        }
        __syncthreads();
    } else {
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (threadIdx.x < 256) hist[threadIdx.x] = 0;
            __syncthreads();
            const unsigned int prefix = ctl[0];
            const unsigned int pmask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int e0 = 0; e0 < E; e0 += FZ_THREADS) {
                const int e = e0 + threadIdx.x;
                const unsigned int kk = e < E ? fkey(sv[e]) : 0u;
                hist_add(hist, e < E && (kk & pmask) == prefix, (kk >> shift) & 255u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                int bin; unsigned int below;
                find_bin(hist, ctl[1], &bin, &below);
                __syncwarp();
                if (threadIdx.x == 0) { ctl[1] -= below; ctl[0] = prefix | ((unsigned int)bin << shift); }
            }
            __syncthreads();
        }
    }
    FZ_S(c, 13);
    const float vq = fkey_inv(ctl[0]);
    __syncthreads();
    // ordered list (flat index ascending): contiguous chunk per thread
    const int per = (E + FZ_THREADS - 1) / FZ_THREADS;
    const int e0 = threadIdx.x * per, e1 = min(E, e0 + per);
    int cnt = 0;
    for (int e = e0; e < e1; ++e) { const float v = sv[e]; cnt += (v >= vq && v != 0.f); }
    int total;
    int o = block_excl_scan512(cnt, &total, s.iscr + 12);
    for (int e = e0; e < e1; ++e) {
        const float v = sv[e];
        if (v >= vq && v != 0.f) { if (o < LIST_CAP) { s.li[o] = e; s.lv[o] = v; } ++o; }
    }
    if (threadIdx.x == 0) s.lc[0] = total;
    __syncthreads();
    FZ_S(c, 14);
    // own rows of the dense tensor + bitmap
    for (int e = c.i0 * FZ_K + threadIdx.x; e < c.i1 * FZ_K; e += FZ_THREADS) {
        const float v = sv[e]; const bool keep = v >= vq;
        xout_g[e] = keep ? v : 0.f;
        bits_g[e] = keep;
    }
    if (c.r == 0) {
        if (threadIdx.x == 0) lcnt_g[0] = total;
        if (threadIdx.x < min(total, LIST_CAP)) { lidx_g[threadIdx.x] = (uint16_t)s.li[threadIdx.x]; lval_g[threadIdx.x] = s.lv[threadIdx.x]; }
    }
    __syncthreads();
    FZ_S(c, 15);
}

}  // namespace fz

// =================================================================================================================================
// forward kernel
// =================================================================================================================================
__global__ void __cluster_dims__(FZ_CL, 1, 1) __launch_bounds__(FZ_THREADS, 1) k_csc_fused_fwd(const FzPlan P, const FzBufs B, const CscDims d) {
    using namespace fz;
    extern __shared__ __align__(16) float fz_smem[];
    const int Lb = d.Lb, cc = d.c, l = d.l;
    const int R = fz_rows(cc);
    Smem s; carve(s, fz_smem, R, Lb);
    Ctx c;
    c.n = blockIdx.x / FZ_CL; c.r = blockIdx.x % FZ_CL; c.g = c.n / d.B; c.gidx = (c.n % d.B) * FZ_CL + c.r; c.ng = d.B * FZ_CL; c.ncl = d.B;
    c.p0 = min(cc, c.r * R); c.p1 = min(cc, c.p0 + R); c.nr = c.p1 - c.p0;
    c.i0 = min(l, c.p0); c.i1 = min(l, c.p1); c.ni = c.i1 - c.i0;
    c.q1 = (c.r == FZ_CL - 1) ? Lb : c.p1;
    c.epoch = 0; c.bar = B.bar + c.g; c.tbar = 0; c.nbar = 0; c.tb1 = c.tb2 = c.tb3 = 0; for (int i = 0; i < 24; ++i) c.sub[i] = 0; c.subt = 0;
    const int nmed = P.npx + 2;
    float* const data = B.data;
    const float* sc = data + P.sc;
    const int64_t nZ = (int64_t)cc * FZ_M, nZY = (int64_t)cc * FZ_M2, nX = (int64_t)l * FZ_K, nS = (int64_t)4 * Lb;
    const int64_t nD = FZ_FLEN * FZ_M, nF = (int64_t)FZ_H * FZ_M2 * FZ_K;
    const float mf = d.mf;
    FZ_TDECL;
    // ---- prologue: filters and bases into shared memory --------------------------------------------------------------------
    stage_f4<8>(reinterpret_cast<float4*>(s.F), reinterpret_cast<const float4*>(data + P.Fe), (int)nF / 4);
    for (int e = threadIdx.x; e < (int)nD; e += FZ_THREADS) s.D[e] = __ldcg(data + P.De + e);
    const int nb_own = min(Lb, c.p1 + 7 + FZ_FL) - c.p0;                       // bases [p0, ...) this CTA ever looks at
    for (int e = threadIdx.x; e < nb_own; e += FZ_THREADS) s.b[e] = B.bases[(size_t)c.n * Lb + c.p0 + e];
    for (int e = threadIdx.x; e < R * FZ_M2; e += FZ_THREADS) s.th[e] = 0.f;
    if (threadIdx.x == 0) s.lc[0] = 0;
    __syncthreads();
    build_Dt(s);
    __syncthreads();

    FZ_T(0);
    // per-sequence views of arena tensors
#define SEQ_Z(off) (data + (off) + (int64_t)c.n * nZ)
#define SEQ_ZY(off) (data + (off) + (int64_t)c.n * nZY)
#define SEQ_X(off) (data + (off) + (int64_t)c.n * nX)
#define SEQ_S(off) (data + (off) + (int64_t)c.n * nS)
#define LCNT_(L) (B.lcnt + (size_t)(L) * d.NS + c.n)
#define LIDX_(L) (B.lidx + ((size_t)(L) * d.NS + c.n) * LIST_CAP)
#define LVAL_(L) (B.lval + ((size_t)(L) * d.NS + c.n) * LIST_CAP)

    // ---- warm-up (model.jl:171-179, 224-232) -------------------------------------------------------------------------------
    {
        const float eta = sc[P.i_eta_w], lam = sc[P.i_lam_w];
        float* zg = SEQ_Z(P.z0); float* yg = SEQ_Z(P.y0);
        for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) {
            const int pl = o / FZ_M, m = o - pl * FZ_M;
            float uf = 0.f, ur = 0.f;
            #pragma unroll
            for (int j = 0; j < FZ_FL; ++j) {
                const int b = s.b[pl + j];
                uf += s.D[(4 * j + b) * FZ_M + m];
                ur += s.D[(4 * (FZ_FL - 1 - j) + 3 - b) * FZ_M + m];
            }
            const float zv = fmaxf(eta * uf - lam * eta, 0.f), yv = fmaxf(eta * ur - lam * eta, 0.f);
            s.z[o] = zv; s.y[o] = yv;
            zg[(size_t)c.p0 * FZ_M + o] = zv; yg[(size_t)c.p0 * FZ_M + o] = yv;
        }
        for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) { s.al[o] = 0.f; s.be[o] = 0.f; }
        for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) s.fx[o] = 0.f;
        __syncthreads();
    }
    FZ_T(1);
    // shared by warm-up and the passes: masked/scaled codes of rows [i0, i1 + 11) -> s.A (optionally d = fx - (zy' - [al be])), corr2d, top-q, tconv
    auto x_chain = [&](int mi, int64_t z_off, int64_t y_off, int64_t med_off, bool with_d, int64_t fx_in, int64_t al_in, int64_t be_in, bool duals_zero,
                       int64_t zy_out, int64_t dd_out, int64_t g_off, int64_t x_in, bool have_prev, int64_t x_out, int64_t bits_off, int xl_out, float om,
                       int64_t fx_out, bool final_mask) {
        const float med = group_median(c, s, B, mi, nmed);
        FZ_T(2);
        if (c.gidx == 0 && threadIdx.x == 0) data[med_off + c.g] = med;
        if (final_mask) {                                      // the mask ADMM_DF works with (model.jl:365): zyF on own rows, nothing else
            float* zo = SEQ_ZY(zy_out);
            for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) {
                const int rl = o / FZ_M2, j = o - rl * FZ_M2, m = j < FZ_M ? j : j - FZ_M;
                const float v = j < FZ_M ? s.z[rl * FZ_M + m] : s.y[rl * FZ_M + m];
                const float zyv = v >= med ? mf * v : 0.f;
                s.zyF[o] = zyv; zo[(size_t)c.p0 * FZ_M2 + o] = zyv;
            }
            __syncthreads();
            return;
        }
        // A tile rows [i0, i1 + 11): own rows from shared memory, the rest recomputed from what the neighbours published
        const int a_hi = min(cc, c.i1 + FZ_H - 1);
        const float* zg = SEQ_Z(z_off); const float* yg = SEQ_Z(y_off);
        const float* fxg = with_d ? SEQ_ZY(fx_in) : nullptr;
        const float* alg = (with_d && !duals_zero) ? SEQ_Z(al_in) : nullptr; const float* beg = (with_d && !duals_zero) ? SEQ_Z(be_in) : nullptr;
        if (c.ni > 0)
        for (int o0 = 0; o0 < (a_hi - c.i0) * FZ_M2; o0 += 4 * FZ_THREADS) {
            // four elements per thread, all their global loads (the 11 halo rows come from what the neighbours published) issued before the first
            // store: the destination is a generic pointer, so the compiler keeps loads behind earlier stores and each element would pay its own L2 round trip
            float v[4], fxv[4], ab[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int o = o0 + u * FZ_THREADS + threadIdx.x;
                v[u] = 0.f; fxv[u] = 0.f; ab[u] = 0.f;
                if (o < (a_hi - c.i0) * FZ_M2) {
                    const int rl = o / FZ_M2, j = o - rl * FZ_M2, row = c.i0 + rl;
                    const bool own = row < c.p1;
                    const int m = j < FZ_M ? j : j - FZ_M;
                    if (own) v[u] = j < FZ_M ? s.z[(row - c.p0) * FZ_M + m] : s.y[(row - c.p0) * FZ_M + m];
                    else v[u] = __ldcg((j < FZ_M ? zg : yg) + (size_t)row * FZ_M + m);
                    if (with_d) {
                        if (own) { fxv[u] = s.fx[(row - c.p0) * FZ_M2 + j]; ab[u] = j < FZ_M ? s.al[(row - c.p0) * FZ_M + m] : s.be[(row - c.p0) * FZ_M + m]; }
                        else { fxv[u] = __ldcg(fxg + (size_t)row * FZ_M2 + j); ab[u] = duals_zero ? 0.f : __ldcg((j < FZ_M ? alg : beg) + (size_t)row * FZ_M + m); }
                    }
                }
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int o = o0 + u * FZ_THREADS + threadIdx.x;
                if (o < (a_hi - c.i0) * FZ_M2) {
                    const int rl = o / FZ_M2, j = o - rl * FZ_M2, row = c.i0 + rl;
                    const bool own = row < c.p1;
                    float zyv = v[u] >= med ? mf * v[u] : 0.f;
                    if (own && zy_out >= 0) SEQ_ZY(zy_out)[(size_t)row * FZ_M2 + j] = zyv;
                    if (with_d) {
                        zyv = fxv[u] - (zyv - ab[u]);
                        if (own && dd_out >= 0) SEQ_ZY(dd_out)[(size_t)row * FZ_M2 + j] = zyv;
                    }
                    s.A[o] = zyv;
                }
            }
        }
        // rows of zy / d that no x row reaches (i >= l) still belong to somebody's tape tensors
        if (c.p1 > max(c.i1 + FZ_H - 1, c.p0) || c.ni == 0) {
            const int r_lo = c.ni > 0 ? max(c.p0, a_hi) : c.p0;
            for (int o = threadIdx.x + (r_lo - c.p0) * FZ_M2; o < c.nr * FZ_M2; o += FZ_THREADS) {
                const int rl = o / FZ_M2, j = o - rl * FZ_M2, m = j < FZ_M ? j : j - FZ_M;
                const float v = j < FZ_M ? s.z[rl * FZ_M + m] : s.y[rl * FZ_M + m];
                float zyv = v >= med ? mf * v : 0.f;
                if (zy_out >= 0) SEQ_ZY(zy_out)[(size_t)(c.p0 + rl) * FZ_M2 + j] = zyv;
                if (with_d) { zyv = s.fx[o] - (zyv - (j < FZ_M ? s.al[rl * FZ_M + m] : s.be[rl * FZ_M + m])); if (dd_out >= 0) SEQ_ZY(dd_out)[(size_t)(c.p0 + rl) * FZ_M2 + j] = zyv; }
            }
        }
        __syncthreads();
        FZ_T(3);
        corr2d_rows(s, s.F, c.ni);
        FZ_T(4);
        float* gg = SEQ_X(g_off);
        for (int o = threadIdx.x; o < c.ni * FZ_K; o += FZ_THREADS) gg[(size_t)c.i0 * FZ_K + o] = s.gout[o];
        cluster_barrier();                                     // the sequence's g is complete
        FZ_T(5);
        topq_all(c, s, gg, have_prev ? SEQ_X(x_in) : nullptr, have_prev, om, l, d.q, SEQ_X(x_out), B.bits + bits_off + (size_t)c.n * nX,
                 LCNT_(xl_out), LIDX_(xl_out), LVAL_(xl_out));
        FZ_T(6);
        tconv_rows(c, s, s.F, SEQ_X(x_out), l, s.fx);
        __syncthreads();
        float* fo = SEQ_ZY(fx_out);
        for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) fo[(size_t)c.p0 * FZ_M2 + o] = s.fx[o];
        FZ_T(7);
    };
    // ---- warm-up x chain (n = -1), the ADMM_XYZ passes (model.jl:256-268), and the batch median + mask ADMM_DF starts with (n = npx): ONE
    //      instance of the chain's code (the kernel is ~18 k SASS instructions; every duplicated phase costs instruction-cache misses on the
    //      critical path of all warps at once) ----------------------------------------------------------------------------------------------
    const int n_end = P.forward_only ? P.npx : P.npx + 1;
    #pragma unroll 1
    for (int n = -1; n < n_end; ++n) {
        const bool wu = n < 0, fin = n == P.npx;
        const FzPass& X = P.px[wu ? 0 : (fin ? P.npx - 1 : n)];
        const float eta = sc[X.i_eta], lam = sc[X.i_lam], rho = sc[X.i_rho], om = sc[X.i_om];
        if (!wu && !fin) {
        // (1) recon of base positions [p0, p1 + 7) from rows [p0 - 7, p1 + 7) of the previous z, y; residual r = recon - S
        {
            const int lo = max(0, c.p0 - 7), hi = min(cc, c.p1 + 7);
            const int q0 = c.p0, q1 = min(Lb, max(c.p1 + 7, c.q1));
            if (c.nr > 0 || c.q1 > c.p0) {
                stage_zy_halo(s, SEQ_Z(X.z_in), SEQ_Z(X.y_in), lo, hi);
                __syncthreads();
                FZ_T(10);
                recon_rows(c, s, lo, hi, q0, q1, -1.f, SEQ_S(X.rec), c.p0, c.q1, Lb);
            }
            __syncthreads();
        }
        FZ_T(8);
        // (2) corr_sig + ISTA step (model.jl:240-244); reads the old z, y, fx, alpha, beta rows held in shared memory
        {
            float* gzg = SEQ_Z(X.gz); float* gyg = SEQ_Z(X.gy); float* zo = SEQ_Z(X.z_out); float* yo = SEQ_Z(X.y_out);
            for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) {
                const int pl = o / FZ_M, m = o - pl * FZ_M;
                float a = 0.f, b = 0.f;
                #pragma unroll 8
                for (int k = 0; k < FZ_FLEN; ++k) { const float r = s.sig[4 * pl + k]; a += r * s.D[k * FZ_M + m]; b += r * s.D[(FZ_FLEN - 1 - k) * FZ_M + m]; }
                const float zv = s.z[o], yv = s.y[o];
                const float lft = s.fx[pl * FZ_M2 + m], rgt = s.fx[pl * FZ_M2 + FZ_M + m];
                const float zn = fmaxf(zv - eta * (a + rho * (zv - lft - s.al[o])) - lam * eta, 0.f);
                const float yn = fmaxf(yv - eta * (b + rho * (yv - rgt - s.be[o])) - lam * eta, 0.f);
                const size_t go = (size_t)c.p0 * FZ_M + o;
                gzg[go] = a; gyg[go] = b; zo[go] = zn; yo[go] = yn;
                s.z[o] = zn; s.y[o] = yn;
            }
            __syncthreads();
        }
        FZ_T(9);
        }
        // (3) mask, d, corr2d, top-q, tconv
        x_chain(n + 1, wu ? P.z0 : X.z_out, wu ? P.y0 : X.y_out, wu ? P.med0 : (fin ? P.medF : X.med), !wu, X.fx_in, X.al_in, X.be_in, n <= 0,
                wu ? P.zy0 : (fin ? P.zyF : (int64_t)-1), wu ? (int64_t)-1 : X.dd, wu ? P.g0 : X.g, X.x_in, !wu, wu ? P.x0 : X.x_out, wu ? P.bits0 : X.bits,
                wu ? P.xl0 : X.xl_out, wu ? sc[P.i_om_w] : -om, wu ? P.fx0 : X.fx_out, fin);
        if (fin) break;
        // (4) duals (model.jl:265-266); the duals after the last pass are never read
        if (!wu && X.al_out >= 0) {
            float* ao = SEQ_Z(X.al_out); float* bo = SEQ_Z(X.be_out);
            for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) {
                const int pl = o / FZ_M, m = o - pl * FZ_M;
                const float an = s.al[o] + s.fx[pl * FZ_M2 + m] - s.z[o], bn = s.be[o] + s.fx[pl * FZ_M2 + FZ_M + m] - s.y[o];
                s.al[o] = an; s.be[o] = bn;
                ao[(size_t)c.p0 * FZ_M + o] = an; bo[(size_t)c.p0 * FZ_M + o] = bn;
            }
        }
        __syncthreads();
    }
    FZ_T(10);
    if (P.forward_only) return;

    // ---- ADMM_DF (model.jl:362-373) ----------------------------------------------------------------------------------------
    const FzPass& XL = P.px[P.npx - 1];
    float* part = B.part + ((size_t)c.g * c.ng + c.gidx) * FZ_PART;
    float* gpart = B.part + (size_t)c.g * c.ng * FZ_PART;
    const int lo7 = max(0, c.p0 - 7), hi7 = min(cc, c.p1 + 7);
    // s.fx holds fx(x, F) of the sequence's final codes with the CURRENT F on own rows: the last pass's x chain left it there, and every DF pass
    // renews it after its F update -- the e of the next pass, theta and the loss all read that one transposed convolution.
    // Iteration n = npd is the loss (model.jl:310-325, with the updated D and F): it shares the recon code with the passes.
    #pragma unroll 1
    for (int n = 0; n <= P.npd; ++n) {
        const bool last = n == P.npd;
        const FzDf& Y = P.df[last ? max(P.npd - 1, 0) : n];
        const float mu = sc[Y.i_mu], kap = sc[Y.i_kap], kaps = sc[Y.i_kaps];
        // D chain: recon with the current D, R = recon + S ('+S': model.jl:282-285; the loss: recon - S), partial 32-lag gradient over own rows
        {
            FZ_S0(c);
            stage_zy_halo(s, SEQ_Z(XL.z_out), SEQ_Z(XL.y_out), lo7, hi7);
            __syncthreads();
            recon_rows(c, s, lo7, hi7, c.p0, min(Lb, max(c.p1 + 7, c.q1)), last ? -1.f : +1.f, SEQ_S(last ? P.recL : Y.rec), c.p0, c.q1, Lb);
            __syncthreads();
            if (last) break;
            // G[tau][m] += z[p][m] R[4p + tau] + y[p][m] R[4p + 31 - tau] over own rows: tile of 4 taus x 2 filters per thread (16-byte signal loads)
            FZ_S(c, 16);
            float* Gp = s.w;                                       // this CTA's partial [32][50]
            for (int o = threadIdx.x; o < (FZ_FLEN / 4) * (FZ_M / 2); o += FZ_THREADS) {
                const int tg = o / (FZ_M / 2), mp = o - tg * (FZ_M / 2);
                float acc[4][2];
                #pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
                for (int pl = 0; pl < c.nr; ++pl) {
                    const float2 z2 = *reinterpret_cast<const float2*>(s.z + pl * FZ_M + 2 * mp), y2 = *reinterpret_cast<const float2*>(s.y + pl * FZ_M + 2 * mp);
                    const float4 rf = *reinterpret_cast<const float4*>(s.sig + 4 * pl + 4 * tg), rr = *reinterpret_cast<const float4*>(s.sig + 4 * pl + FZ_FLEN - 4 - 4 * tg);
                    const float f[4] = {rf.x, rf.y, rf.z, rf.w}, rv[4] = {rr.w, rr.z, rr.y, rr.x};
                    #pragma unroll
                    for (int i = 0; i < 4; ++i) { acc[i][0] += z2.x * f[i] + y2.x * rv[i]; acc[i][1] += z2.y * f[i] + y2.y * rv[i]; }
                }
                #pragma unroll
                for (int i = 0; i < 4; ++i) { Gp[(4 * tg + i) * FZ_M + 2 * mp] = acc[i][0]; Gp[(4 * tg + i) * FZ_M + 2 * mp + 1] = acc[i][1]; }
            }
            FZ_S(c, 17);
            // sum over the CTAs of the cluster through distributed shared memory (rank order), one slice of 1600 / FZ_CL entries per CTA; the 6 cluster
            // sums go through global memory
            cluster_barrier();
            {
                cg::cluster_group cl = cg::this_cluster();
                float* cpart = B.part + ((size_t)c.g * c.ng + (c.n % d.B) * FZ_CL) * FZ_PART;      // the cluster's slot (its rank-0 CTA's)
                const int per = (int)nD / FZ_CL;
                for (int o = threadIdx.x; o < per; o += FZ_THREADS) {
                    const int e = c.r * per + o;
                    float acc = 0.f;
                    #pragma unroll
                    for (int q = 0; q < FZ_CL; ++q) acc += cl.map_shared_rank(Gp, q)[e];
                    cpart[e] = acc;
                }
            }
            cluster_barrier();                                     // remote reads of this CTA's partial are done: s.w may be reused
            FZ_S(c, 18);
        }
        FZ_T(11);
        // F chain: e = fx(x, F) - (zyF + theta) on own rows (model.jl:294)
        {
            float* eg = SEQ_ZY(Y.e);
            for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) eg[(size_t)c.p0 * FZ_M2 + o] = s.fx[o] - s.zyF[o] - s.th[o];
        }
        FZ_T(12);
        group_barrier(c);                                      // partial D gradients and e are published
        // D update, redundantly in every CTA: G = sum of the group's partials in CTA order; D <- D exp(-mu G), renormalised (model.jl:287-288)
        {
            float* Gs = s.w;
            for (int o = threadIdx.x; o < (int)nD; o += FZ_THREADS) {
                float v[8];
                #pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = q < d.B ? __ldcg(gpart + (size_t)q * FZ_CL * FZ_PART + o) : 0.f;       // all loads in flight together
                float acc = 0.f;
                #pragma unroll
                for (int q = 0; q < 8; ++q) acc += v[q];
                for (int q = 8; q < d.B; ++q) acc += __ldcg(gpart + (size_t)q * FZ_CL * FZ_PART + o);
                Gs[o] = acc;
                if (c.gidx == 0) data[Y.Gm + (int64_t)c.g * nD + o] = acc;
            }
            __syncthreads();
            for (int o = threadIdx.x; o < FZ_FL * FZ_M; o += FZ_THREADS) {
                const int m = o % FZ_M, j = o / FZ_M;
                float u[4], sum = 0.f;
                #pragma unroll
                for (int a = 0; a < 4; ++a) { const int e = (4 * j + a) * FZ_M + m; u[a] = s.D[e] * expf(-mu * Gs[e]); sum += u[a]; }
                #pragma unroll
                for (int a = 0; a < 4; ++a) { const int e = (4 * j + a) * FZ_M + m; const float v = u[a] / sum; s.D[e] = v; if (c.gidx == 0) data[Y.Dn + (int64_t)c.g * nD + e] = v; }
            }
            __syncthreads();
            build_Dt(s);
            __syncthreads();
        }
        FZ_T(13);
        // F gradient + update for syntax filter k = gidx (model.jl:292-308): Fg[a][j][k] = sum_b sum_i e[b][a+i][j] x[b][i][k]
        if (c.gidx < FZ_K) {
            const int k = c.gidx;
            int* sel_b = reinterpret_cast<int*>(s.A); int* sel_i = sel_b + 512; float* sel_v = reinterpret_cast<float*>(sel_i + 512);
            int* cnt_p = s.iscr;
            if (threadIdx.x == 0) { cnt_p[0] = 0; cnt_p[1] = 0; }
            __syncthreads();
            // gather the group's codes of filter k in (sequence, position) order: thread = (sequence b, slot q)
            {
                const int b = threadIdx.x / LIST_CAP, q = threadIdx.x - b * LIST_CAP;
                bool hit = false; int e = 0; float v = 0.f;
                if (b < d.B) {
                    const int nn = c.g * d.B + b;
                    const int cn = __ldcg(B.lcnt + (size_t)XL.xl_out * d.NS + nn);
                    if (cn > LIST_CAP) cnt_p[1] = 1;
                    else if (q < cn) { e = __ldcg(B.lidx + ((size_t)XL.xl_out * d.NS + nn) * LIST_CAP + q); hit = (e % FZ_K) == k; if (hit) v = __ldcg(B.lval + ((size_t)XL.xl_out * d.NS + nn) * LIST_CAP + q); }
                }
                int total;
                const int o = block_excl_scan512(hit ? 1 : 0, &total, s.iscr + 8);
                if (hit && o < 512) { sel_b[o] = b; sel_i[o] = e / FZ_K; sel_v[o] = v; }
                if (threadIdx.x == 0) cnt_p[0] = total;
            }
            __syncthreads();
            const bool dense = cnt_p[1] != 0 || cnt_p[0] > 512 || d.B * LIST_CAP > FZ_THREADS;
            const int ns = cnt_p[0];
            const float* Fin = data + Y.F_in + (int64_t)c.g * Y.F_in_gs;
            float ss = 0.f;
            float uu[3], gg[3];
            #pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int o = threadIdx.x + it * FZ_THREADS;           // o = a*100 + j
                uu[it] = 0.f; gg[it] = 0.f;
                if (o < FZ_H * FZ_M2) {
                    const int a = o / FZ_M2, j = o - a * FZ_M2;
                    float acc = 0.f;
                    if (!dense) {
                        for (int q0 = 0; q0 < ns; q0 += 8) {                 // eight independent loads per trip (one L2 round trip, not eight)
                            float ev[8];
                            #pragma unroll
                            for (int u = 0; u < 8; ++u) { const int q = min(q0 + u, ns - 1); ev[u] = __ldcg(data + Y.e + ((int64_t)(c.g * d.B + sel_b[q]) * cc + a + sel_i[q]) * FZ_M2 + j); }
                            #pragma unroll
                            for (int u = 0; u < 8; ++u) if (q0 + u < ns) acc += sel_v[q0 + u] * ev[u];
                        }
                    } else {
                        for (int b = 0; b < d.B; ++b)
                            for (int i = 0; i < l; ++i) {
                                const float xv = __ldcg(data + XL.x_out + ((int64_t)(c.g * d.B + b) * l + i) * FZ_K + k);
                                if (xv != 0.f) acc += xv * __ldcg(data + Y.e + ((int64_t)(c.g * d.B + b) * cc + a + i) * FZ_M2 + j);
                            }
                    }
                    gg[it] = acc;
                    const float u = fmaxf(__ldcg(Fin + (size_t)o * FZ_K + k) - kap * acc - kap * kaps, 0.f);
                    uu[it] = u; ss += u * u;
                }
            }
            ss = block_sum512(ss, s.red);
            const float nn = sqrtf(ss);
            if (threadIdx.x == 0) data[Y.nrm + (int64_t)c.g * FZ_K + k] = nn;
            #pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int o = threadIdx.x + it * FZ_THREADS;
                if (o < FZ_H * FZ_M2) {
                    data[Y.Fg + (int64_t)c.g * nF + (size_t)o * FZ_K + k] = gg[it];
                    data[Y.Fn + (int64_t)c.g * nF + (size_t)o * FZ_K + k] = uu[it] / nn;
                }
            }
        }
        FZ_T(14);
        group_barrier(c);                                      // the updated F is published
        stage_f4<8>(reinterpret_cast<float4*>(s.F), reinterpret_cast<const float4*>(data + Y.Fn + (int64_t)c.g * nF), (int)nF / 4);
        __syncthreads();
        // fx(x, F_new): read by theta <- theta + fx - zyF (model.jl:370, only needed by the next pass), by the next pass's e and by the loss
        tconv_rows(c, s, s.F, SEQ_X(XL.x_out), l, s.fx);
        __syncthreads();
        if (Y.has_theta_out) {
            float* tg = SEQ_ZY(Y.thn);
            for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) { const float t = s.fx[o] - s.zyF[o] + s.th[o]; s.th[o] = t; tg[(size_t)c.p0 * FZ_M2 + o] = t; }
            __syncthreads();
        }
    }
    FZ_T(15);
    FZ_S0(c);
    // ---- loss (model.jl:310-325) with the updated D, F -----------------------------------------------------------------------
    {   // the residual of the loss's recon is in s.sig (last trip of the loop above), fx(x, F) with the final F in s.fx
        FZ_S(c, 19);
        float a = 0.f, b = 0.f;
        for (int t = threadIdx.x; t < 4 * (c.q1 - c.p0); t += FZ_THREADS) { const float r = s.sig[t]; a += r * r; }
        float* fg = SEQ_ZY(P.fxL);
        for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) { const float v = s.fx[o]; fg[(size_t)c.p0 * FZ_M2 + o] = v; const float dlt = v - s.zyF[o]; b += dlt * dlt; }
        FZ_S(c, 20);
        a = block_sum512(a, s.red);
        b = block_sum512(b, s.red);
        FZ_S(c, 21);
        if (threadIdx.x == 0) { part[0] = a; part[1] = b; }
        group_barrier(c);
        FZ_S(c, 22);
        if (c.gidx == 0) {
            float* la = s.w; float* lb = s.w + 256;                 // one load per thread (a serial loop would pay an L2 round trip per term)
            if (threadIdx.x < c.ng) { la[threadIdx.x] = __ldcg(gpart + (size_t)threadIdx.x * FZ_PART); lb[threadIdx.x] = __ldcg(gpart + (size_t)threadIdx.x * FZ_PART + 1); }
            __syncthreads();
            if (threadIdx.x == 0) {
                float sa = 0.f, sb = 0.f;
                for (int q = 0; q < c.ng; ++q) { sa += la[q]; sb += lb[q]; }      // CTA order: deterministic
                data[P.loss + c.g * 3 + 0] = (sa + sb) / (float)d.B; data[P.loss + c.g * 3 + 1] = sa / (float)d.B; data[P.loss + c.g * 3 + 2] = sb / (float)d.B;
            }
        }
    }
#ifdef FZ_PROFILE
    FZ_T(11);
    {   // pure barrier latency: 16 back-to-back barriers with no work in between
        const long long t0 = clock64();
        for (int i = 0; i < 16; ++i) group_barrier(c);
        const long long t1 = clock64();
        for (int i = 0; i < 16; ++i) cluster_barrier();
        const long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) printf("[fz] back-to-back: group barrier %lld clk, cluster barrier %lld clk\n", (t1 - t0) / 16, (t2 - t1) / 16);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const char* nm[16] = {"prologue", "warm_zy", "median", "A tile", "corr2d", "g+cluster bar", "topq", "tconv+store", "recon", "corr_sig+zy", "stage zy halo", "DF D-chain+loss", "DF e", "barA+D update", "F update", "barB+F load+theta"};
        long long tot = 0; for (int i = 0; i < 16; ++i) tot += fz_acc[i];
        for (int i = 0; i < 16; ++i) printf("[fz] %-20s %9lld clk %5.1f%%\n", nm[i], fz_acc[i], 100.0 * fz_acc[i] / tot);
        printf("[fz] median: hist %lld push %lld bar1 %lld pull %lld find %lld cand %lld bar2 %lld radix %lld v2 %lld | topq: values %lld hist %lld find %lld select %lld list %lld out %lld\n",
               c.sub[0], c.sub[1], c.sub[2], c.sub[3], c.sub[4], c.sub[5], c.sub[6], c.sub[7], c.sub[8], c.sub[10], c.sub[11], c.sub[12], c.sub[13], c.sub[14], c.sub[15]);
        printf("[fz] loss: recon %lld tconv+sum %lld blocksums %lld barrier %lld\n", c.sub[19], c.sub[20], c.sub[21], c.sub[22]);
        printf("[fz] DF D-chain: stage+recon %lld Gp %lld cluster reduce %lld\n", c.sub[16], c.sub[17], c.sub[18]);
        printf("[fz] total %lld clk; %d group barriers, %lld clk inside them (thread 0): fence %lld, arrive+poll %lld, fence %lld\n", tot, c.nbar, c.tbar, c.tb1, c.tb2, c.tb3);
    }
#endif
#undef SEQ_Z
#undef SEQ_ZY
#undef SEQ_X
#undef SEQ_S
#undef LCNT_
#undef LIDX_
#undef LVAL_
}

// =================================================================================================================================
// reverse pass of the ADMM_XYZ passes (the bulk of the step: 6 of the 9 unrolled passes) as one persistent kernel
// =================================================================================================================================
// Hand-derived adjoints of one_forward_step_XYZ (model.jl:256-268), same cluster-per-sequence layout as the forward kernel.  The
// adjoint of a pass's outputs (z+, y+, x+, fx+, alpha+, beta+) is turned into the adjoint of its inputs; masks, top-q supports and the
// zero duals are constants exactly where the reference uses @ignore.  Three cluster barriers per pass (halo exchanges of dfx+, of the
// sparse dx values, of dgz/dgy) and no barrier across sequences: gradients of the shared parameters (D, F, the per-pass scalars) are
// accumulated per CTA and summed once at the end in CTA order (deterministic).  The adjoint of x is only ever needed on the support the
// top-q kept (model.jl:190 masks everything else), so the adjoint of fx = x (*) F is a handful of 1200-term dot products per sequence
// instead of the dense contraction.
#define FZ_KCAP 256               // kept entries per sequence the kernel handles (q = 32 plus ties); more raises the error flag

struct FzBwd {                    // extra buffers of the reverse pass
    float* grad;                  // the tape's adjoint arena (same offsets as data)
    float* dFp;                   // [NS][h*K][2M] per-sequence partial of dF, layout [a][k][j] (row t = a*K + k written by the cluster's CTA t % FZ_CL)
    float* xch;                   // [NS][FZ_KCAP] exchange of the sparse dx values inside a cluster
    float* gsum;                  // [G][nF + nD + 64] group sums: dF (layout [a][j][k]), dD, dsc
    unsigned int* err;            // set when a sequence has more than FZ_KCAP kept entries / an overlong code list
};

namespace fz {

#define FZ_NTGT ((FZ_H * FZ_K + FZ_CL - 1) / FZ_CL)      // dF targets (a, k) a CTA owns: t = a*K + k with t % FZ_CL == rank
struct SmemB {
    float *F, *D, *Dt, *Dr;
    float *dFt;                                         // [FZ_NTGT][2M] this CTA's partial of dF for the targets it owns
    float *dz, *dy, *dal, *dbe, *gzs, *gys;             // [R][50]
    float *dfx;                                         // [R][100]
    float *A;                                           // work tile (own rows [R][100] / halo staging)
    float *sig, *sig2;                                  // drec and r = rec - S over base positions [p0, p1 + 7)
    float *w;                                           // row lists of tconv
    float *dDp, *dscp;                                  // partial dD [32][50], partial scalar gradients [64]
    int *kl; float *kv;                                 // kept entries (flat index) and their values
    int *klp;                                           // kept entries of the pass handled before (whose d x values this pass consumes and clears)
    int *li2; float *lv2;                               // code list of x+ (from the tape)
    float* red; int* iscr; uint8_t* b;
};
__host__ __device__ inline size_t fzb_smem_bytes(int Lb) {
    const int c = Lb - FZ_FL + 1, R = fz_rows(c);
    size_t f = (size_t)FZ_H * FZ_M2 * FZ_K + 3 * FZ_FLEN * FZ_M + (size_t)FZ_NTGT * FZ_M2;
    f += (size_t)R * (FZ_M * 6 + FZ_M2);
    const size_t A = (size_t)R * FZ_M2, A2 = (size_t)2 * (R + 14) * FZ_M;
    f += (A > A2 ? A : A2);
    f += (size_t)8 * (R + 8);
    f += 32 + 64 * LIST_CAP;                              // row lists
    f += FZ_FLEN * FZ_M + 64;
    f += 3 * FZ_KCAP + 2 * LIST_CAP + 64;
    return f * 4 + (size_t)(R + 16) + 64;
}
__device__ __forceinline__ void carve_b(SmemB& s, float* base, int R, int Lb) {
    float* p = base;
    s.F = p; p += FZ_H * FZ_M2 * FZ_K;
    s.D = p; p += FZ_FLEN * FZ_M; s.Dt = p; p += FZ_FLEN * FZ_M; s.Dr = p; p += FZ_FLEN * FZ_M;
    s.dFt = p; p += FZ_NTGT * FZ_M2;
    s.dz = p; p += R * FZ_M; s.dy = p; p += R * FZ_M; s.dal = p; p += R * FZ_M; s.dbe = p; p += R * FZ_M;
    s.gzs = p; p += R * FZ_M; s.gys = p; p += R * FZ_M;
    s.dfx = p; p += R * FZ_M2;
    const size_t A = (size_t)R * FZ_M2, A2 = (size_t)2 * (R + 14) * FZ_M;
    s.A = p; p += (A > A2 ? A : A2);
    s.sig = p; p += 4 * (R + 8); s.sig2 = p; p += 4 * (R + 8);
    s.w = p; p += 32 + 64 * LIST_CAP;
    s.dDp = p; p += FZ_FLEN * FZ_M; s.dscp = p; p += 64;
    s.kl = reinterpret_cast<int*>(p); p += FZ_KCAP; s.kv = p; p += FZ_KCAP; s.klp = reinterpret_cast<int*>(p); p += FZ_KCAP;
    s.li2 = reinterpret_cast<int*>(p); p += LIST_CAP; s.lv2 = p; p += LIST_CAP;
    s.red = p; p += 32; s.iscr = reinterpret_cast<int*>(p); p += 32;
    s.b = reinterpret_cast<uint8_t*>(p);
}

// F_gradient (model.jl:292-302) of one sequence, balanced over its cluster whatever rows the codes sit in: dF[a][k][:] += sum over the list
// entries (i, k, v) of v * A[i + a][:].  CTA r of the cluster owns the targets (a, k) with (a*K + k) % FZ_CL == r; one warp per owned target
// walks the list in order (fixed summation order), reads the rows of A from global memory (published before the preceding barrier) and
// does ONE update of the CTA's shared-memory partial.  No two warps share a target, no atomics.
__device__ void fgrad_targets(const Ctx& c, const float* Ag /* global [c][2M] of the sequence */, const int* li, const float* lv, int cnt, float* dFt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = c.r + FZ_CL * warp; t < FZ_H * FZ_K; t += FZ_CL * (FZ_THREADS / 32)) {
        const int a = t / FZ_K, k = t - a * FZ_K;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        bool any = false;
        for (int q0 = 0; q0 < cnt; q0 += 32) {
            const int q = q0 + lane;
            int e = -1;
            if (q < cnt) e = li[q];
            const bool hit = e >= 0 && (e % FZ_K) == k;
            unsigned mk = __ballot_sync(FULLMASK, hit);
            while (mk) {
                // up to four hits per trip: all their row loads are in flight together (an L2 round trip per hit would dominate otherwise)
                float v[4]; const float* row[4];
                #pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                    if (mk) {
                        const int src = __ffs(mk) - 1; mk &= mk - 1;
                        const int ee = __shfl_sync(FULLMASK, e, src);
                        v[hh] = lv[q0 + src]; row[hh] = Ag + (size_t)(ee / FZ_K + a) * FZ_M2;
                    } else { v[hh] = 0.f; row[hh] = Ag; }
                }
                float x0[4], x1[4], x2[4], x3[4];
                #pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                    x0[hh] = __ldcg(row[hh] + lane); x1[hh] = __ldcg(row[hh] + 32 + lane); x2[hh] = __ldcg(row[hh] + 64 + lane);
                    x3[hh] = lane < FZ_M2 - 96 ? __ldcg(row[hh] + 96 + lane) : 0.f;
                }
                #pragma unroll
                for (int hh = 0; hh < 4; ++hh) { acc0 += v[hh] * x0[hh]; acc1 += v[hh] * x1[hh]; acc2 += v[hh] * x2[hh]; acc3 += v[hh] * x3[hh]; }
                any = true;
            }
        }
        if (any) {
            float* dst = dFt + (size_t)(t / FZ_CL) * FZ_M2;
            dst[lane] += acc0; dst[32 + lane] += acc1; dst[64 + lane] += acc2;
            if (lane < FZ_M2 - 96) dst[96 + lane] += acc3;
        }
    }
    __syncthreads();
}

}  // namespace fz

__global__ void __cluster_dims__(FZ_CL, 1, 1) __launch_bounds__(FZ_THREADS, 1) k_csc_fused_bwd_xyz(const FzPlan P, const FzBufs B, const FzBwd W, const CscDims d) {
    using namespace fz;
    extern __shared__ __align__(16) float fz_smem[];
    const int Lb = d.Lb, cc = d.c, l = d.l;
    const int R = fz_rows(cc);
    SmemB s; carve_b(s, fz_smem, R, Lb);
    Ctx c;
    c.n = blockIdx.x / FZ_CL; c.r = blockIdx.x % FZ_CL; c.g = c.n / d.B; c.gidx = (c.n % d.B) * FZ_CL + c.r; c.ng = d.B * FZ_CL; c.ncl = d.B;
    c.p0 = min(cc, c.r * R); c.p1 = min(cc, c.p0 + R); c.nr = c.p1 - c.p0;
    c.i0 = min(l, c.p0); c.i1 = min(l, c.p1); c.ni = c.i1 - c.i0;
    c.q1 = (c.r == FZ_CL - 1) ? Lb : c.p1;
    c.epoch = 0; c.bar = B.bar + c.g; c.tbar = 0; c.nbar = 0; c.tb1 = c.tb2 = c.tb3 = 0; for (int i = 0; i < 24; ++i) c.sub[i] = 0; c.subt = 0;
    float* const data = B.data; float* const grad = W.grad;
    const float* sc = data + P.sc;
    const int64_t nZ = (int64_t)cc * FZ_M, nZY = (int64_t)cc * FZ_M2, nX = (int64_t)l * FZ_K, nS = (int64_t)4 * Lb;
    const int nD = FZ_FLEN * FZ_M, nF = FZ_H * FZ_M2 * FZ_K;
    const float mf = d.mf;
    const int E = l * FZ_K;
    FZ_TDECL;
    // Smem view that the shared device functions (recon_rows, tconv_list) expect
    Smem sv; sv.F = s.F; sv.D = s.D; sv.Dt = s.Dt; sv.Dr = s.Dr; sv.A = s.A; sv.sig = s.sig; sv.w = s.w; sv.b = s.b; sv.red = s.red; sv.iscr = s.iscr;
    sv.z = sv.y = sv.fx = sv.al = sv.be = sv.th = sv.zyF = sv.gout = nullptr; sv.li = s.kl; sv.lv = s.kv; sv.lc = nullptr;
    float* dFp = s.dFt;                                           // partial dF of the targets this CTA owns, in shared memory
#define SEQ(off, per) ((off) + (int64_t)c.n * (per))
    // ---- prologue ----------------------------------------------------------------------------------------------------------------
    stage_f4<8>(reinterpret_cast<float4*>(s.F), reinterpret_cast<const float4*>(data + P.Fe), nF / 4);
    for (int e = threadIdx.x; e < nD; e += FZ_THREADS) { s.D[e] = __ldcg(data + P.De + e); s.dDp[e] = 0.f; }
    if (threadIdx.x < 64) s.dscp[threadIdx.x] = 0.f;
    for (int e = threadIdx.x; e < FZ_NTGT * FZ_M2; e += FZ_THREADS) s.dFt[e] = 0.f;
    const int nb_own = min(Lb, c.p1 + 7 + FZ_FL) - c.p0;
    for (int e = threadIdx.x; e < nb_own; e += FZ_THREADS) s.b[e] = B.bases[(size_t)c.n * Lb + c.p0 + e];
    const FzPass& XL = P.px[P.npx - 1];
    {   // adjoints handed over by the ops after the passes (DF, loss): dz, dy of the final codes, dx of the final x (dense)
        const float* gz_ = grad + SEQ(XL.z_out, nZ) + (size_t)c.p0 * FZ_M; const float* gy_ = grad + SEQ(XL.y_out, nZ) + (size_t)c.p0 * FZ_M;
        for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) { s.dz[o] = __ldcg(gz_ + o); s.dy[o] = __ldcg(gy_ + o); s.dal[o] = 0.f; s.dbe[o] = 0.f; }
        for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) s.dfx[o] = 0.f;
    }
    __syncthreads();
    build_Dt(sv);
    __syncthreads();

    FZ_T(0);
    int cnt_prev = 0;
    for (int n = P.npx - 1; n >= 0; --n) {
        const FzPass& X = P.px[n];
        const float eta = sc[X.i_eta], lam = sc[X.i_lam], rho = sc[X.i_rho], om = sc[X.i_om];
        // kept entries of x+ (the top-q bitmap of this pass), ordered by flat index; redundantly in every CTA of the cluster
        int cnt;
        if (n != P.npx - 1) { if ((int)threadIdx.x < cnt_prev) s.klp[threadIdx.x] = s.kl[threadIdx.x]; __syncthreads(); }
        {
            const uint32_t* bw = reinterpret_cast<const uint32_t*>(B.bits + X.bits + (size_t)c.n * nX);      // 4 flags per word (E is a multiple of 4)
            const int nw = E >> 2;
            const int per = (nw + FZ_THREADS - 1) / FZ_THREADS;
            const int w0 = threadIdx.x * per, w1 = min(nw, w0 + per);
            int k = 0;
            for (int w = w0; w < w1; ++w) { const uint32_t v = bw[w]; k += ((v & 0xffu) != 0) + ((v & 0xff00u) != 0) + ((v & 0xff0000u) != 0) + ((v >> 24) != 0); }
            int total;
            int o = block_excl_scan512(k, &total, s.iscr + 8);
            for (int w = w0; w < w1; ++w) {
                const uint32_t v = bw[w];
                #pragma unroll
                for (int bb = 0; bb < 4; ++bb) if ((v >> (8 * bb)) & 0xffu) { if (o < FZ_KCAP) s.kl[o] = 4 * w + bb; ++o; }
            }
            if (total > FZ_KCAP) { if (threadIdx.x == 0) atomicOr(W.err, 1u); total = FZ_KCAP; }
            cnt = total;
        }
        // the d x+ values this pass is about to read were written on the support of the pass handled before it (for the last pass: by the
        // DF reverse kernel, on this pass's own support); they are cleared after the reads, so the adjoint arena never needs a memset
        if (n == P.npx - 1) { __syncthreads(); if ((int)threadIdx.x < cnt) s.klp[threadIdx.x] = s.kl[threadIdx.x]; cnt_prev = cnt; }
        FZ_T(1);
        // (i') duals: alpha+ = alpha + fx+_l - z+, beta+ likewise (model.jl:265-266); not computed by the last pass
        if (X.al_out >= 0)
            for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) {
                const int pl = o / FZ_M, m = o - pl * FZ_M;
                const float da = s.dal[o], db = s.dbe[o];
                s.dfx[pl * FZ_M2 + m] += da; s.dfx[pl * FZ_M2 + FZ_M + m] += db; s.dz[o] -= da; s.dy[o] -= db;
            }
        __syncthreads();
        float* dfx_g = grad + SEQ(X.fx_out, nZY);                 // scratch: the tape's slot of d fx+ (unused otherwise)
        for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) dfx_g[(size_t)c.p0 * FZ_M2 + o] = s.dfx[o];
        // code list of x+ from the tape (for the F gradient of fx+ = x+ (*) F)
        int cnt2 = __ldcg(B.lcnt + (size_t)X.xl_out * d.NS + c.n);
        if (cnt2 > LIST_CAP) { if (threadIdx.x == 0) atomicOr(W.err, 2u); cnt2 = LIST_CAP; }
        if (threadIdx.x < cnt2) { s.li2[threadIdx.x] = __ldcg(B.lidx + ((size_t)X.xl_out * d.NS + c.n) * LIST_CAP + threadIdx.x); s.lv2[threadIdx.x] = __ldcg(B.lval + ((size_t)X.xl_out * d.NS + c.n) * LIST_CAP + threadIdx.x); }
        FZ_T(2);
        cluster_barrier();                                        // #1: every CTA's rows of d fx+ are published
        FZ_T(3);
        // (h') adjoint of fx+ = x+ (*) F on the kept support: entry q is taken by CTA q % FZ_CL, one warp per entry
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            float* xch = W.xch + (size_t)c.n * FZ_KCAP;
            for (int q = c.r + FZ_CL * warp; q < cnt; q += FZ_CL * (FZ_THREADS / 32)) {
                const int e = s.kl[q], i = e / FZ_K, k = e - i * FZ_K;
                const float* rows = grad + SEQ(X.fx_out, nZY) + (size_t)i * FZ_M2;        // rows i .. i+11 are contiguous: 1200 floats
                float acc = 0.f;
                float xr[38];
                #pragma unroll
                for (int u = 0; u < 38; ++u) { const int t = lane + 32 * u; xr[u] = t < FZ_H * FZ_M2 ? __ldcg(rows + t) : 0.f; }      // 1200 = 37.5 x 32
                #pragma unroll
                for (int u = 0; u < 38; ++u) { const int t = lane + 32 * u; if (t < FZ_H * FZ_M2) acc += xr[u] * s.F[(size_t)t * FZ_K + k]; }
                acc = warp_sum(acc);
                if (lane == 0) xch[q] = acc + __ldcg(grad + SEQ(X.x_out, nX) + e);          // + what later consumers of x+ already left there
            }
        }
        FZ_T(4);
        fgrad_targets(c, grad + SEQ(X.fx_out, nZY), s.li2, s.lv2, cnt2, dFp);
        FZ_T(5);
        cluster_barrier();                                        // #2: the sparse d x+ values are published
        if (c.r == 1 % FZ_CL && (int)threadIdx.x < cnt_prev) grad[SEQ(X.x_out, nX) + s.klp[threadIdx.x]] = 0.f;      // consumed by every CTA before the barrier
        cnt_prev = cnt;
        FZ_T(6);
        // (g') top-q adjoint (model.jl:190-192, 252-253): gr = d x+ on the kept support; d x = gr, d g = -omega gr, d omega = -sum gr g
        {
            const float* xch = W.xch + (size_t)c.n * FZ_KCAP;
            const float* gg = data + SEQ(X.g, nX);
            float som = 0.f;
            float* gxin = grad + SEQ(X.x_in, nX);                  // d x of the pass's input: zero except on the kept support
            if (threadIdx.x < cnt) {
                const float gr = __ldcg(xch + threadIdx.x);
                const int e = s.kl[threadIdx.x];
                if (c.r == 0) gxin[e] = gr;
                som = -gr * __ldcg(gg + e);
                s.kv[threadIdx.x] = -om * gr;                      // the d g list shares the kept entries
            }
            som = block_sum512(som, s.red);
            if (threadIdx.x == 0 && c.r == 0) s.dscp[X.i_om] += som;
            __syncthreads();
        }
        FZ_T(7);
        // (f') adjoint of g = corr2d(dd, F): d dd = dg (*) F on own rows; dF += F_gradient(dd, dg)
        tconv_list(c, sv, s.F, s.kl, s.kv, cnt, FZ_KCAP, nullptr, l, s.A);
        FZ_T(8);
        fgrad_targets(c, data + SEQ(X.dd, nZY), s.kl, s.kv, cnt, dFp);
        FZ_T(9);
        // (e', d', c') d_build, mask, ISTA step (model.jl:248-250, 206-210, 240-244) on own rows
        {
            const float med = __ldcg(data + X.med + c.g);
            const size_t ro = (size_t)c.p0 * FZ_M;
            const float* zo = data + SEQ(X.z_out, nZ) + ro; const float* yo = data + SEQ(X.y_out, nZ) + ro;
            const float* zi = data + SEQ(X.z_in, nZ) + ro; const float* yi = data + SEQ(X.y_in, nZ) + ro;
            const float* gzd = data + SEQ(X.gz, nZ) + ro; const float* gyd = data + SEQ(X.gy, nZ) + ro;
            const float* fxi = data + SEQ(X.fx_in, nZY) + (size_t)c.p0 * FZ_M2;
            const float* ali = n > 0 ? data + SEQ(X.al_in, nZ) + ro : nullptr; const float* bei = n > 0 ? data + SEQ(X.be_in, nZ) + ro : nullptr;
            float* dgz_g = grad + SEQ(X.gz, nZ) + ro; float* dgy_g = grad + SEQ(X.gy, nZ) + ro;
            float s_eta = 0.f, s_lam = 0.f, s_rho = 0.f;
            for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) {
                const int pl = o / FZ_M, m = o - pl * FZ_M;
                const float zov = __ldcg(zo + o), yov = __ldcg(yo + o), ziv = __ldcg(zi + o), yiv = __ldcg(yi + o);
                const float gzv = __ldcg(gzd + o), gyv = __ldcg(gyd + o);
                const float lft = __ldcg(fxi + pl * FZ_M2 + m), rgt = __ldcg(fxi + pl * FZ_M2 + FZ_M + m);
                const float av = ali ? __ldcg(ali + o) : 0.f, bv = bei ? __ldcg(bei + o) : 0.f;
                const float dl = s.A[pl * FZ_M2 + m], dr = s.A[pl * FZ_M2 + FZ_M + m];          // d dd
                float dzo = s.dz[o], dyo = s.dy[o];
                if (zov >= med) dzo += mf * (-dl);
                if (yov >= med) dyo += mf * (-dr);
                const float tz = zov > 0.f ? dzo : 0.f, ty = yov > 0.f ? dyo : 0.f;
                const float ez = ziv - lft - av, ey = yiv - rgt - bv;
                s.dz[o] = tz * (1.f - eta * rho); s.dy[o] = ty * (1.f - eta * rho);
                const float dgz = -eta * tz, dgy = -eta * ty;
                s.gzs[o] = dgz; s.gys[o] = dgy; dgz_g[o] = dgz; dgy_g[o] = dgy;
                s.dfx[pl * FZ_M2 + m] = dl + eta * rho * tz; s.dfx[pl * FZ_M2 + FZ_M + m] = dr + eta * rho * ty;
                s.dal[o] += dl + eta * rho * tz; s.dbe[o] += dr + eta * rho * ty;
                s_eta += tz * (-(gzv + rho * ez) - lam) + ty * (-(gyv + rho * ey) - lam);
                s_lam += -eta * (tz + ty);
                s_rho += -eta * (tz * ez + ty * ey);
            }
            s_eta = block_sum512(s_eta, s.red); s_lam = block_sum512(s_lam, s.red); s_rho = block_sum512(s_rho, s.red);
            if (threadIdx.x == 0) { s.dscp[X.i_eta] += s_eta; s.dscp[X.i_lam] += s_lam; s.dscp[X.i_rho] += s_rho; }
        }
        FZ_T(10);
        cluster_barrier();                                        // #3: every CTA's rows of dgz, dgy are published
        FZ_T(11);
        // (b') d rec = recon(dgz, dgy; D) over base positions [p0, p1 + 7) (halo rows recomputed from the neighbours' dgz, dgy); r = rec - S
        {
            const int lo = max(0, c.p0 - 7), hi = min(cc, c.p1 + 7);
            const int q0 = c.p0, q1 = min(Lb, c.p1 + 7);
            if (c.nr > 0) {
                stage_zy_halo(sv, grad + SEQ(X.gz, nZ), grad + SEQ(X.gy, nZ), lo, hi);
                const float* recg = data + SEQ(X.rec, nS);
                for (int t = threadIdx.x; t < 4 * (q1 - q0); t += FZ_THREADS) {
                    const int q = q0 + (t >> 2);
                    s.sig2[t] = __ldcg(recg + 4 * q0 + t) - (s.b[q - c.p0] == (t & 3) ? 1.f : 0.f);
                }
                __syncthreads();
                recon_rows(c, sv, lo, hi, q0, q1, 0.f, nullptr, 0, 0, Lb);
            }
            __syncthreads();
            // the pass's input codes (own rows) for dgrad(z, y; d rec): the halo staging area is free again
            const float* zi = data + SEQ(X.z_in, nZ) + (size_t)c.p0 * FZ_M; const float* yi = data + SEQ(X.y_in, nZ) + (size_t)c.p0 * FZ_M;
            for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) { s.A[o] = __ldcg(zi + o); s.A[R * FZ_M + o] = __ldcg(yi + o); }
            __syncthreads();
        }
        FZ_T(12);
        // dD += dgrad(dgz, dgy; r) + dgrad(z, y; d rec)  (32-lag gradient, model.jl:270-290 form) over own rows; then
        // (a') dz += corr_sig(d rec; D), dy likewise
        for (int o = threadIdx.x; o < (FZ_FLEN / 4) * (FZ_M / 2); o += FZ_THREADS) {          // tile: taus 4tg..4tg+3 x filters 2mp, 2mp+1
            const int tg = o / (FZ_M / 2), mp = o - tg * (FZ_M / 2);
            float acc[4][2];
            #pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
            for (int pl = 0; pl < c.nr; ++pl) {
                const float2 gz2 = *reinterpret_cast<const float2*>(s.gzs + pl * FZ_M + 2 * mp), gy2 = *reinterpret_cast<const float2*>(s.gys + pl * FZ_M + 2 * mp);
                const float2 z2 = *reinterpret_cast<const float2*>(s.A + pl * FZ_M + 2 * mp), y2 = *reinterpret_cast<const float2*>(s.A + R * FZ_M + pl * FZ_M + 2 * mp);
                const float4 rf = *reinterpret_cast<const float4*>(s.sig2 + 4 * pl + 4 * tg), rr = *reinterpret_cast<const float4*>(s.sig2 + 4 * pl + FZ_FLEN - 4 - 4 * tg);
                const float4 df = *reinterpret_cast<const float4*>(s.sig + 4 * pl + 4 * tg), dr = *reinterpret_cast<const float4*>(s.sig + 4 * pl + FZ_FLEN - 4 - 4 * tg);
                // tau = 4tg + i reads r[4p + tau] (forward: .x .y .z .w) and r[4p + 31 - tau] (reversed block: .w .z .y .x)
                const float f[4] = {rf.x, rf.y, rf.z, rf.w}, rv[4] = {rr.w, rr.z, rr.y, rr.x}, g[4] = {df.x, df.y, df.z, df.w}, gv[4] = {dr.w, dr.z, dr.y, dr.x};
                #pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] += gz2.x * f[i] + gy2.x * rv[i] + z2.x * g[i] + y2.x * gv[i];
                    acc[i][1] += gz2.y * f[i] + gy2.y * rv[i] + z2.y * g[i] + y2.y * gv[i];
                }
            }
            #pragma unroll
            for (int i = 0; i < 4; ++i) { s.dDp[(4 * tg + i) * FZ_M + 2 * mp] += acc[i][0]; s.dDp[(4 * tg + i) * FZ_M + 2 * mp + 1] += acc[i][1]; }
        }
        FZ_T(13);
        for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) {
            const int pl = o / FZ_M, m = o - pl * FZ_M;
            float a = 0.f, b = 0.f;
            #pragma unroll 8
            for (int k = 0; k < FZ_FLEN; ++k) { const float r = s.sig[4 * pl + k]; a += r * s.D[k * FZ_M + m]; b += r * s.D[(FZ_FLEN - 1 - k) * FZ_M + m]; }
            s.dz[o] += a; s.dy[o] += b;
        }
        __syncthreads();
    }
    FZ_T(14);
    // ---- reverse pass of the warm-up (model.jl:224-232): fx0 = tconv(x0, F), x0 = top-q(om_w g0), g0 = corr2d(zy0', F), zy0' = mask(z0, y0),
    //      (z0, y0) = relu(eta_w corr(S, D) - lam_w eta_w); the warm-up scalars are not trained ------------------------------------------------
    {
        int cnt;
        if ((int)threadIdx.x < cnt_prev) s.klp[threadIdx.x] = s.kl[threadIdx.x];
        __syncthreads();
        {   // kept entries of x0
            const uint32_t* bw = reinterpret_cast<const uint32_t*>(B.bits + P.bits0 + (size_t)c.n * nX);
            const int nw = E >> 2;
            const int per = (nw + FZ_THREADS - 1) / FZ_THREADS;
            const int w0 = threadIdx.x * per, w1 = min(nw, w0 + per);
            int k = 0;
            for (int w = w0; w < w1; ++w) { const uint32_t v = bw[w]; k += ((v & 0xffu) != 0) + ((v & 0xff00u) != 0) + ((v & 0xff0000u) != 0) + ((v >> 24) != 0); }
            int total;
            int o = block_excl_scan512(k, &total, s.iscr + 8);
            for (int w = w0; w < w1; ++w) {
                const uint32_t v = bw[w];
                #pragma unroll
                for (int bb = 0; bb < 4; ++bb) if ((v >> (8 * bb)) & 0xffu) { if (o < FZ_KCAP) s.kl[o] = 4 * w + bb; ++o; }
            }
            if (total > FZ_KCAP) { if (threadIdx.x == 0) atomicOr(W.err, 1u); total = FZ_KCAP; }
            cnt = total;
        }
        float* dfx_g = grad + SEQ(P.fx0, nZY);                    // scratch: the tape's slot of d fx0
        for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) dfx_g[(size_t)c.p0 * FZ_M2 + o] = s.dfx[o];
        int cnt2 = __ldcg(B.lcnt + (size_t)P.xl0 * d.NS + c.n);
        if (cnt2 > LIST_CAP) { if (threadIdx.x == 0) atomicOr(W.err, 2u); cnt2 = LIST_CAP; }
        if (threadIdx.x < cnt2) { s.li2[threadIdx.x] = __ldcg(B.lidx + ((size_t)P.xl0 * d.NS + c.n) * LIST_CAP + threadIdx.x); s.lv2[threadIdx.x] = __ldcg(B.lval + ((size_t)P.xl0 * d.NS + c.n) * LIST_CAP + threadIdx.x); }
        cluster_barrier();                                        // d fx0 rows (and the d x0 values pass 0 left) are published
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            float* xch = W.xch + (size_t)c.n * FZ_KCAP;
            for (int q = c.r + FZ_CL * warp; q < cnt; q += FZ_CL * (FZ_THREADS / 32)) {
                const int e = s.kl[q], i = e / FZ_K, k = e - i * FZ_K;
                const float* rows = dfx_g + (size_t)i * FZ_M2;
                float acc = 0.f;
                float xr[38];
                #pragma unroll
                for (int u = 0; u < 38; ++u) { const int t = lane + 32 * u; xr[u] = t < FZ_H * FZ_M2 ? __ldcg(rows + t) : 0.f; }
                #pragma unroll
                for (int u = 0; u < 38; ++u) { const int t = lane + 32 * u; if (t < FZ_H * FZ_M2) acc += xr[u] * s.F[(size_t)t * FZ_K + k]; }
                acc = warp_sum(acc);
                if (lane == 0) xch[q] = acc + __ldcg(grad + SEQ(P.x0, nX) + e);
            }
        }
        fgrad_targets(c, dfx_g, s.li2, s.lv2, cnt2, dFp);          // d F += fgrad(d fx0, x0)
        cluster_barrier();                                        // the sparse d x0 values are published
        if (c.r == 1 % FZ_CL && (int)threadIdx.x < cnt_prev) grad[SEQ(P.x0, nX) + s.klp[threadIdx.x]] = 0.f;           // what pass 0 left there is consumed
        {
            const float om_w = sc[P.i_om_w];
            const float* xch = W.xch + (size_t)c.n * FZ_KCAP;
            if (threadIdx.x < cnt) s.kv[threadIdx.x] = om_w * __ldcg(xch + threadIdx.x);      // d g0 on the kept support
            __syncthreads();
        }
        tconv_list(c, sv, s.F, s.kl, s.kv, cnt, FZ_KCAP, nullptr, l, s.A);                   // d zy0' = tconv(d g0; F) on own rows
        fgrad_targets(c, data + SEQ(P.zy0, nZY), s.kl, s.kv, cnt, dFp);                     // d F += fgrad(zy0', d g0)
        {
            const float med = __ldcg(data + P.med0 + c.g), eta_w = sc[P.i_eta_w];
            const float* z0 = data + SEQ(P.z0, nZ) + (size_t)c.p0 * FZ_M; const float* y0 = data + SEQ(P.y0, nZ) + (size_t)c.p0 * FZ_M;
            for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) {
                const int pl = o / FZ_M, m = o - pl * FZ_M;
                const float zv = __ldcg(z0 + o), yv = __ldcg(y0 + o);
                float dzv = s.dz[o], dyv = s.dy[o];
                if (zv >= med) dzv += mf * s.A[pl * FZ_M2 + m];
                if (yv >= med) dyv += mf * s.A[pl * FZ_M2 + FZ_M + m];
                s.gzs[o] = zv > 0.f ? eta_w * dzv : 0.f;
                s.gys[o] = yv > 0.f ? eta_w * dyv : 0.f;
            }
            __syncthreads();
            // d D[4j + a][m] += sum over own rows p with base[p + j] == a of gz[p][m], + those with base[p + 7 - j] == 3 - a of gy[p][m]
            // (the reverse-complement filter reads D[4(7 - j') + 3 - b]); one thread per (j, m), fixed order
            for (int o = threadIdx.x; o < FZ_FL * FZ_M; o += FZ_THREADS) {
                const int j = o / FZ_M, m = o - j * FZ_M;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                for (int pl = 0; pl < c.nr; ++pl) {
                    const int bz = s.b[pl + j], by = 3 - (int)s.b[pl + FZ_FL - 1 - j];
                    const float gz = s.gzs[pl * FZ_M + m], gy = s.gys[pl * FZ_M + m];
                    a0 += (bz == 0 ? gz : 0.f) + (by == 0 ? gy : 0.f); a1 += (bz == 1 ? gz : 0.f) + (by == 1 ? gy : 0.f);
                    a2 += (bz == 2 ? gz : 0.f) + (by == 2 ? gy : 0.f); a3 += (bz == 3 ? gz : 0.f) + (by == 3 ? gy : 0.f);
                }
                s.dDp[(4 * j + 0) * FZ_M + m] += a0; s.dDp[(4 * j + 1) * FZ_M + m] += a1; s.dDp[(4 * j + 2) * FZ_M + m] += a2; s.dDp[(4 * j + 3) * FZ_M + m] += a3;
            }
            __syncthreads();
        }
    }
    // ---- gradients of the shared parameters: per-CTA partials -> group sums in CTA order ------------------------------------------------
    float* part = B.part + ((size_t)c.g * c.ng + c.gidx) * FZ_PART;
    for (int o = threadIdx.x; o < nD; o += FZ_THREADS) part[o] = s.dDp[o];
    if (threadIdx.x < 64) part[nD + threadIdx.x] = s.dscp[threadIdx.x];
    for (int o = threadIdx.x; o < FZ_NTGT * FZ_M2; o += FZ_THREADS) {
        const int lt = o / FZ_M2, t = lt * FZ_CL + c.r;
        if (t < FZ_H * FZ_K) W.dFp[((size_t)c.n * (FZ_H * FZ_K) + t) * FZ_M2 + (o - lt * FZ_M2)] = s.dFt[o];
    }
    group_barrier(c);
    {
        const float* gpart = B.part + (size_t)c.g * c.ng * FZ_PART;
        const float* gF = W.dFp + (size_t)c.g * d.B * nF;            // [B sequences][a*K + k][j]
        float* gs = W.gsum + (size_t)c.g * (nF + nD + 64);
        const int per = (nF + c.ng - 1) / c.ng;
        for (int po = c.gidx * per + threadIdx.x; po < min(nF, (c.gidx + 1) * per); po += FZ_THREADS) {      // po = (a*K + k)*2M + j: coalesced reads
            float acc = 0.f;
            for (int q = 0; q < d.B; ++q) acc += __ldcg(gF + (size_t)q * nF + po);
            const int j = po % FZ_M2, ak = po / FZ_M2, k = ak % FZ_K, a = ak / FZ_K;
            gs[(a * FZ_M2 + j) * FZ_K + k] = acc;
        }
        const int perd = (nD + 64 + c.ng - 1) / c.ng;
        for (int o = c.gidx * perd + threadIdx.x; o < min(nD + 64, (c.gidx + 1) * perd); o += FZ_THREADS) {
            float acc = 0.f;
            #pragma unroll 8
            for (int q = 0; q < c.ng; ++q) acc += __ldcg(gpart + (size_t)q * FZ_PART + o);
            gs[nF + o] = acc;
        }
    }
#ifdef FZ_PROFILE
    FZ_T(15);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const char* nm[16] = {"prologue", "kept list", "duals+publish+list", "barrier 1", "sparse dx dots", "fgrad(dfx,x)", "barrier 2", "topq adjoint", "tconv(dg)", "fgrad(dd,dg)", "elementwise", "barrier 3", "recon(dgz)+r", "dgrad", "corr_sig(drec)", "epilogue"};
        long long tot = 0; for (int i = 0; i < 16; ++i) tot += fz_acc[i];
        for (int i = 0; i < 16; ++i) printf("[fzb] %-20s %9lld clk %5.1f%%\n", nm[i], fz_acc[i], 100.0 * fz_acc[i] / tot);
        printf("[fzb] total %lld clk\n", tot);
    }
    if (blockIdx.x < 16 && threadIdx.x == 0)
        printf("[fzb-cta %2d] dots %6lld fgrad1 %6lld b2 %6lld | topq %6lld tconv %6lld fgrad2 %6lld elem %6lld b3 %6lld | recon %6lld dgrad %6lld b1 %6lld kept %6lld duals %6lld\n", (int)blockIdx.x,
               fz_acc[4], fz_acc[5], fz_acc[6], fz_acc[7], fz_acc[8], fz_acc[9], fz_acc[10], fz_acc[11], fz_acc[12], fz_acc[13], fz_acc[3], fz_acc[1], fz_acc[2]);
#endif
#undef SEQ
}

// =================================================================================================================================
// reverse pass of everything after the ADMM_XYZ passes: loss, the ADMM_DF passes, the final mask (model.jl:310-325, 362-373, 206-210)
// =================================================================================================================================
// Same cluster-per-sequence layout.  Inputs are the tape tensors the forward kernel left (final z, y, code list, zy', and per DF pass
// rec, G, D+, e, Fg, F+, ||.||); outputs are d z, d y of the final codes (own rows), d x on the support the last top-q kept (all the
// XYZ reverse pass reads), and the group sums of d D0, d F0, d mu/kappa/kappa_s.  Per pass, in reverse: the adjoints of the two
// GROUP-level updates need batch sums, so a pass costs two group barriers like its forward:
//   B1: per-sequence partials of d F+ (owned targets, via dFp) and of d D+ (cluster sums, via part) are published
//       -> d_update adjoint redundantly in every CTA (d D, d G, d mu); f_update adjoint of filter k in CTA k (d F kept there, d Fg published)
//   B2: d Fg is published -> per sequence: u = tconv(x; d Fg) (adjoint of e), d x += corr2d(e; d Fg) + corr2d(w; F) on the kept support,
//       d F += fgrad(w, x);  d rec = recon(z, y; d G), d z,y += corr_sig(rec + S; d G) + corr_sig(d rec; D), d D += dgrad(z, y; d rec).
// fx(x, F_n) enters e_n and theta_n with opposite signs, so the adjoint w_n of that product is u_0 for the first pass, the adjoint of
// theta_{n+1} for the middle passes and exactly zero for the last one (the tape adds and subtracts the same term there).
namespace fz {

struct SmemDf {
    float *Dc, *dG, *Dt, *Dr, *dDg, *Gp;                // current D, d G, its two tap-major forms, group-level part of d D, this CTA's dgrad partial
    float *sig, *sigB;                                  // recon output (d rec) / signal loaded from the tape (rec + S, or the loss residual)
    float *zy;                                          // z | y rows [p0 - 7, p1 + 7) of the final codes (constant)
    float *dzyF, *dth, *u;                              // [R][100]: adjoint of zy', of theta, tconv output
    float *w;                                           // row lists of tconv
    float *dFt;                                         // [FZ_NTGT][100] partial d F of the owned targets
    float *duK;                                         // [h*2M] d F (group-level part) of the filter this CTA owns
    float *dxv; int* kl; int* li2; float* lv2;          // d x on the kept support, kept entries, code list
    float *dscp, *red; int* iscr;
    float *dz, *dy;                                     // [R][50]
    uint8_t* b;
};
__host__ __device__ inline size_t fzd_smem_bytes(int Lb) {
    const int c = Lb - FZ_FL + 1, R = fz_rows(c);
    size_t f = (size_t)6 * FZ_FLEN * FZ_M + (size_t)8 * (R + 8) + (size_t)2 * (R + 14) * FZ_M + (size_t)3 * R * FZ_M2;
    f += 32 + 64 * LIST_CAP;
    f += (size_t)FZ_NTGT * FZ_M2 + (size_t)FZ_H * FZ_M2 + 2 * FZ_KCAP + 2 * LIST_CAP + 64 + 32 + 32;
    f += (size_t)2 * R * FZ_M + 8;
    return f * 4 + (size_t)(R + 16) + 64;
}
__device__ __forceinline__ void carve_df(SmemDf& s, float* base, int R) {
    float* p = base;
    s.Dc = p; p += FZ_FLEN * FZ_M; s.dG = p; p += FZ_FLEN * FZ_M; s.Dt = p; p += FZ_FLEN * FZ_M; s.Dr = p; p += FZ_FLEN * FZ_M;
    s.dDg = p; p += FZ_FLEN * FZ_M; s.Gp = p; p += FZ_FLEN * FZ_M;
    s.sig = p; p += 4 * (R + 8); s.sigB = p; p += 4 * (R + 8);
    s.zy = p; p += 2 * (R + 14) * FZ_M;
    s.dzyF = p; p += R * FZ_M2; s.dth = p; p += R * FZ_M2; s.u = p; p += R * FZ_M2;
    s.w = p; p += 32 + 64 * LIST_CAP;
    s.dFt = p; p += FZ_NTGT * FZ_M2;
    s.duK = p; p += FZ_H * FZ_M2;
    s.dxv = p; p += FZ_KCAP; s.kl = reinterpret_cast<int*>(p); p += FZ_KCAP;
    s.li2 = reinterpret_cast<int*>(p); p += LIST_CAP; s.lv2 = p; p += LIST_CAP;
    s.dscp = p; p += 64; s.red = p; p += 32; s.iscr = reinterpret_cast<int*>(p); p += 32;
    s.dz = p; p += R * FZ_M; s.dy = p; p += R * FZ_M;
    p += 8;
    s.b = reinterpret_cast<uint8_t*>(p);
}

// kept entries (flat index, ascending) of a top-q bitmap [E bytes]; the same list in every CTA of the cluster.  Returns the count (capped).
__device__ int kept_from_bits(const uint8_t* bits, int E, int* kl, int* iscr, unsigned int* err) {
    const uint32_t* bw = reinterpret_cast<const uint32_t*>(bits);
    const int nw = E >> 2;
    const int per = (nw + FZ_THREADS - 1) / FZ_THREADS;
    const int w0 = threadIdx.x * per, w1 = min(nw, w0 + per);
    int k = 0;
    for (int w = w0; w < w1; ++w) { const uint32_t v = bw[w]; k += ((v & 0xffu) != 0) + ((v & 0xff00u) != 0) + ((v & 0xff0000u) != 0) + ((v >> 24) != 0); }
    int total;
    int o = block_excl_scan512(k, &total, iscr);
    for (int w = w0; w < w1; ++w) {
        const uint32_t v = bw[w];
        #pragma unroll
        for (int bb = 0; bb < 4; ++bb) if ((v >> (8 * bb)) & 0xffu) { if (o < FZ_KCAP) kl[o] = 4 * w + bb; ++o; }
    }
    if (total > FZ_KCAP) { if (threadIdx.x == 0) atomicOr(err, 1u); total = FZ_KCAP; }
    __syncthreads();
    return total;
}

// d x[e] += sum_{t < h*2M} rows[i*2M + t] * Fm[t*K + k] for the kept entries e = i*K + k this CTA owns (entry q -> CTA q % FZ_CL, one warp each)
__device__ void dots_kept(const Ctx& c, const float* rows_g, const float* Fm, const int* kl, int cnt, float* dxv, int st = FZ_K, int sk = 1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = c.r + FZ_CL * warp; q < cnt; q += FZ_CL * (FZ_THREADS / 32)) {
        const int e = kl[q], i = e / FZ_K, k = e - i * FZ_K;
        const float* rows = rows_g + (size_t)i * FZ_M2;
        float xr[38], fr[38];
        #pragma unroll
        for (int u = 0; u < 38; ++u) { const int t = lane + 32 * u; xr[u] = t < FZ_H * FZ_M2 ? __ldcg(rows + t) : 0.f; }
        #pragma unroll
        for (int u = 0; u < 38; ++u) { const int t = lane + 32 * u; fr[u] = t < FZ_H * FZ_M2 ? __ldcg(Fm + (size_t)t * st + (size_t)k * sk) : 0.f; }
        float acc = 0.f;
        #pragma unroll
        for (int u = 0; u < 38; ++u) acc += xr[u] * fr[u];
        acc = warp_sum(acc);
        if (lane == 0) dxv[q] += acc;
    }
    __syncthreads();
}

// own rows: d z,y += corr_sig(sig1; f1) (+ corr_sig(sig2; f2)) and this CTA's partial Gp = dgrad(z, y; sigG); signals cover base positions [p0, p1 + 7)
__device__ void d_data_step(const Ctx& c, const float* zown, const float* yown, const float* sig1, const float* f1, const float* sig2, const float* f2,
                            const float* sigG, float* Gp, float* dz, float* dy) {
    const int nt = (FZ_FLEN / 4) * (FZ_M / 2);                     // 200 tiles of 4 taus x 2 filters
    if ((int)threadIdx.x < nt) {
        const int o = threadIdx.x, tg = o / (FZ_M / 2), mp = o - tg * (FZ_M / 2);
        float acc[4][2];
        #pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
        for (int pl = 0; pl < c.nr; ++pl) {
            const float2 z2 = *reinterpret_cast<const float2*>(zown + pl * FZ_M + 2 * mp), y2 = *reinterpret_cast<const float2*>(yown + pl * FZ_M + 2 * mp);
            const float4 rf = *reinterpret_cast<const float4*>(sigG + 4 * pl + 4 * tg), rr = *reinterpret_cast<const float4*>(sigG + 4 * pl + FZ_FLEN - 4 - 4 * tg);
            const float f[4] = {rf.x, rf.y, rf.z, rf.w}, rv[4] = {rr.w, rr.z, rr.y, rr.x};
            #pragma unroll
            for (int i = 0; i < 4; ++i) { acc[i][0] += z2.x * f[i] + y2.x * rv[i]; acc[i][1] += z2.y * f[i] + y2.y * rv[i]; }
        }
        #pragma unroll
        for (int i = 0; i < 4; ++i) { Gp[(4 * tg + i) * FZ_M + 2 * mp] = acc[i][0]; Gp[(4 * tg + i) * FZ_M + 2 * mp + 1] = acc[i][1]; }
    } else {
        for (int o = threadIdx.x - nt; o < c.nr * FZ_M; o += FZ_THREADS - nt) {
            const int pl = o / FZ_M, m = o - pl * FZ_M;
            float a = 0.f, b = 0.f;
            #pragma unroll 8
            for (int k = 0; k < FZ_FLEN; ++k) { const float r = sig1[4 * pl + k]; a += r * f1[k * FZ_M + m]; b += r * f1[(FZ_FLEN - 1 - k) * FZ_M + m]; }
            if (sig2) {
                #pragma unroll 8
                for (int k = 0; k < FZ_FLEN; ++k) { const float r = sig2[4 * pl + k]; a += r * f2[k * FZ_M + m]; b += r * f2[(FZ_FLEN - 1 - k) * FZ_M + m]; }
            }
            dz[o] += a; dy[o] += b;
        }
    }
    __syncthreads();
}

}  // namespace fz

__global__ void __cluster_dims__(FZ_CL, 1, 1) __launch_bounds__(FZ_THREADS, 1) k_csc_fused_bwd_df(const FzPlan P, const FzBufs B, const FzBwd W, float* __restrict__ gsum2, const CscDims d) {
    using namespace fz;
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) float fz_smem[];
    const int Lb = d.Lb, cc = d.c, l = d.l;
    const int R = fz_rows(cc);
    SmemDf s; carve_df(s, fz_smem, R);
    Ctx c;
    c.n = blockIdx.x / FZ_CL; c.r = blockIdx.x % FZ_CL; c.g = c.n / d.B; c.gidx = (c.n % d.B) * FZ_CL + c.r; c.ng = d.B * FZ_CL; c.ncl = d.B;
    c.p0 = min(cc, c.r * R); c.p1 = min(cc, c.p0 + R); c.nr = c.p1 - c.p0;
    c.i0 = min(l, c.p0); c.i1 = min(l, c.p1); c.ni = c.i1 - c.i0;
    c.q1 = (c.r == FZ_CL - 1) ? Lb : c.p1;
    c.epoch = 0; c.bar = B.bar + c.g; c.tbar = 0; c.nbar = 0; c.tb1 = c.tb2 = c.tb3 = 0; for (int i = 0; i < 24; ++i) c.sub[i] = 0; c.subt = 0;
    float* const data = B.data; float* const grad = W.grad;
    const float* sc = data + P.sc;
    const int64_t nZ = (int64_t)cc * FZ_M, nZY = (int64_t)cc * FZ_M2, nX = (int64_t)l * FZ_K, nS = (int64_t)4 * Lb;
    const int nD = FZ_FLEN * FZ_M, nF = FZ_H * FZ_M2 * FZ_K, HJ = FZ_H * FZ_M2;
    const float mf = d.mf;
    FZ_TDECL;
    Smem sv; sv.F = nullptr; sv.D = s.dG; sv.Dt = s.Dt; sv.Dr = s.Dr; sv.A = s.zy; sv.sig = s.sig; sv.w = s.w; sv.b = s.b; sv.red = s.red; sv.iscr = s.iscr;
    sv.z = sv.y = sv.fx = sv.al = sv.be = sv.th = sv.zyF = sv.gout = nullptr; sv.li = s.li2; sv.lv = s.lv2; sv.lc = nullptr;
#define SEQ(off, per) ((off) + (int64_t)c.n * (per))
    const FzPass& XL = P.px[P.npx - 1];
    const int lo7 = max(0, c.p0 - 7), hi7 = min(cc, c.p1 + 7);
    const int q0 = c.p0, q1 = min(Lb, c.p1 + 7);
    const int nq4 = c.nr > 0 ? 4 * (q1 - q0) : 0;
    const float* zown = s.zy + (c.p0 - lo7) * FZ_M; const float* yown = s.zy + (hi7 - lo7) * FZ_M + (c.p0 - lo7) * FZ_M;
    float* part = B.part + ((size_t)c.g * c.ng + c.gidx) * FZ_PART;
    const float* gpart = B.part + (size_t)c.g * c.ng * FZ_PART;
    float* cpart = B.part + ((size_t)c.g * c.ng + (c.n % d.B) * FZ_CL) * FZ_PART;            // the cluster's slot (its rank-0 CTA's)
    cg::cluster_group cl = cg::this_cluster();
    // sum of the cluster's dgrad partials through distributed shared memory (rank order) -> the cluster's slot; call between two cluster barriers
    auto cluster_reduce_Gp = [&]() {
        const int per = nD / FZ_CL;
        for (int o = threadIdx.x; o < per; o += FZ_THREADS) {
            const int e = c.r * per + o;
            float acc = 0.f;
            #pragma unroll
            for (int q = 0; q < FZ_CL; ++q) acc += cl.map_shared_rank(s.Gp, q)[e];
            cpart[e] = acc;
        }
    };
    auto publish_dFt = [&]() {
        for (int o = threadIdx.x; o < FZ_NTGT * FZ_M2; o += FZ_THREADS) {
            const int lt = o / FZ_M2, t = lt * FZ_CL + c.r;
            if (t < FZ_H * FZ_K) W.dFp[((size_t)c.n * (FZ_H * FZ_K) + t) * FZ_M2 + (o - lt * FZ_M2)] = s.dFt[o];
        }
    };
    // ---- prologue ----------------------------------------------------------------------------------------------------------------
    {
        const int nb_own = min(Lb, c.p1 + 7 + FZ_FL) - c.p0;
        for (int e = threadIdx.x; e < nb_own; e += FZ_THREADS) s.b[e] = B.bases[(size_t)c.n * Lb + c.p0 + e];
        for (int o = threadIdx.x; o < R * FZ_M; o += FZ_THREADS) { s.dz[o] = 0.f; s.dy[o] = 0.f; }
        for (int o = threadIdx.x; o < R * FZ_M2; o += FZ_THREADS) { s.dzyF[o] = 0.f; s.dth[o] = 0.f; }
        for (int e = threadIdx.x; e < nD; e += FZ_THREADS) s.dDg[e] = 0.f;
        for (int e = threadIdx.x; e < HJ; e += FZ_THREADS) s.duK[e] = 0.f;
        if (threadIdx.x < 64) s.dscp[threadIdx.x] = 0.f;
        if (threadIdx.x < FZ_KCAP) s.dxv[threadIdx.x] = 0.f;
        if (hi7 > lo7) stage_zy_halo(sv, data + SEQ(XL.z_out, nZ), data + SEQ(XL.y_out, nZ), lo7, hi7);
    }
    int cnt2 = __ldcg(B.lcnt + (size_t)XL.xl_out * d.NS + c.n);
    if (cnt2 > LIST_CAP) { if (threadIdx.x == 0) atomicOr(W.err, 2u); cnt2 = LIST_CAP; }
    if ((int)threadIdx.x < cnt2) { s.li2[threadIdx.x] = __ldcg(B.lidx + ((size_t)XL.xl_out * d.NS + c.n) * LIST_CAP + threadIdx.x); s.lv2[threadIdx.x] = __ldcg(B.lval + ((size_t)XL.xl_out * d.NS + c.n) * LIST_CAP + threadIdx.x); }
    const int cnt = kept_from_bits(B.bits + XL.bits + (size_t)c.n * nX, l * FZ_K, s.kl, s.iscr + 8, W.err);
    FZ_T(0);
    // ---- loss (model.jl:310-325): d recL = sw (recL - S), d fxL = sw (fxL - zy'), d zy' = -d fxL ------------------------------------------
    {
        const float sw = 2.f / ((float)d.G * (float)d.B);
        const FzDf& YL = P.df[P.npd - 1];
        for (int e = threadIdx.x; e < nD; e += FZ_THREADS) s.Dc[e] = __ldcg(data + YL.Dn + (int64_t)c.g * nD + e);
        const float* recg = data + SEQ(P.recL, nS);
        for (int t = threadIdx.x; t < nq4; t += FZ_THREADS) {
            const int q = q0 + (t >> 2);
            s.sigB[t] = sw * (__ldcg(recg + 4 * q0 + t) - (s.b[q - c.p0] == (t & 3) ? 1.f : 0.f));
        }
        const float* fg = data + SEQ(P.fxL, nZY) + (size_t)c.p0 * FZ_M2; const float* zyg = data + SEQ(P.zyF, nZY) + (size_t)c.p0 * FZ_M2;
        float* scr = grad + SEQ(P.fxL, nZY) + (size_t)c.p0 * FZ_M2;
        for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) { const float v = sw * (__ldcg(fg + o) - __ldcg(zyg + o)); s.dzyF[o] = -v; scr[o] = v; }
        for (int o = threadIdx.x; o < FZ_NTGT * FZ_M2; o += FZ_THREADS) s.dFt[o] = 0.f;
        __syncthreads();
        d_data_step(c, zown, yown, s.sigB, s.Dc, nullptr, nullptr, s.sigB, s.Gp, s.dz, s.dy);
        cluster_barrier();                                        // d fxL rows and the dgrad partials of the cluster are complete
        cluster_reduce_Gp();
        dots_kept(c, grad + SEQ(P.fxL, nZY), data + YL.Fn + (int64_t)c.g * nF, s.kl, cnt, s.dxv);
        fgrad_targets(c, grad + SEQ(P.fxL, nZY), s.li2, s.lv2, cnt2, s.dFt);
        publish_dFt();
        cluster_barrier();                                        // remote reads of Gp are done
    }
    FZ_T(1);
    // ---- ADMM_DF passes in reverse ---------------------------------------------------------------------------------------------------
    for (int n = P.npd - 1; n >= 0; --n) {
        const FzDf& Y = P.df[n];
        const float mu = sc[Y.i_mu], kap = sc[Y.i_kap], kaps = sc[Y.i_kaps];
        const bool has_partial = (n + 1 == P.npd) || (n + 1 <= P.npd - 2);      // some sequence-level term reached d F_{n+1}
        const bool has_w = (n == 0) || (n <= P.npd - 2);
        group_barrier(c);                                         // B1
        FZ_T(2);
        // d_update adjoint (model.jl:287-288), redundantly: d D+ = group-level part + the clusters' data terms
        {
            for (int o = threadIdx.x; o < nD; o += FZ_THREADS) {
                float v[8];
                #pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = q < d.B ? __ldcg(gpart + (size_t)q * FZ_CL * FZ_PART + o) : 0.f;
                float acc = s.dDg[o];
                #pragma unroll
                for (int q = 0; q < 8; ++q) acc += v[q];
                s.dDg[o] = acc;
                s.Dc[o] = __ldcg(data + Y.D_in + (int64_t)c.g * Y.D_in_gs + o);
            }
            __syncthreads();
            float s_mu = 0.f;
            for (int o = threadIdx.x; o < FZ_FL * FZ_M; o += FZ_THREADS) {
                const int m = o % FZ_M, j = o / FZ_M;
                float Gv[4], ex[4], uu[4], dDn[4], ssum = 0.f, dot = 0.f;
                #pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int e = (4 * j + a) * FZ_M + m;
                    Gv[a] = __ldcg(data + Y.Gm + (int64_t)c.g * nD + e); ex[a] = expf(-mu * Gv[a]); uu[a] = s.Dc[e] * ex[a]; ssum += uu[a];
                    dDn[a] = s.dDg[e]; dot += dDn[a] * __ldcg(data + Y.Dn + (int64_t)c.g * nD + e);
                }
                #pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int e = (4 * j + a) * FZ_M + m;
                    const float du = (dDn[a] - dot) / ssum;
                    s.dDg[e] = du * ex[a]; s.dG[e] = du * uu[a] * (-mu); s_mu += du * uu[a] * (-Gv[a]);
                }
            }
            s_mu = block_sum512(s_mu, s.red);
            if (threadIdx.x == 0 && c.gidx == 0) s.dscp[Y.i_mu] += s_mu;
            __syncthreads();
            build_Dt(sv);                                         // tap-major forms of d G for recon
        }
        FZ_T(3);
        // f_update adjoint (model.jl:304-308) of filter k = gidx: d F+ = what this CTA kept + the sequences' partials; d Fg published
        float* dFg_g = grad + Y.Fg + (int64_t)c.g * nF;           // scratch: the tape's adjoint slot of Fg
        if (c.gidx < FZ_K) {
            const int k = c.gidx;
            float dFv[3], fnv[3], dot = 0.f;
            #pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int e = threadIdx.x + it * FZ_THREADS;
                dFv[it] = 0.f; fnv[it] = 0.f;
                if (e < HJ) {
                    float acc = s.duK[e];
                    if (has_partial) {
                        const int a = e / FZ_M2, j = e - a * FZ_M2;
                        float v[8];
                        #pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = q < d.B ? __ldcg(W.dFp + ((size_t)(c.g * d.B + q) * (FZ_H * FZ_K) + a * FZ_K + k) * FZ_M2 + j) : 0.f;
                        #pragma unroll
                        for (int q = 0; q < 8; ++q) acc += v[q];
                    }
                    dFv[it] = acc; fnv[it] = __ldcg(data + Y.Fn + (int64_t)c.g * nF + (size_t)e * FZ_K + k);
                    dot += acc * fnv[it];
                }
            }
            dot = block_sum512(dot, s.red);
            const float nn = __ldcg(data + Y.nrm + (int64_t)c.g * FZ_K + k);
            float s_kap = 0.f, s_kaps = 0.f;
            #pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int e = threadIdx.x + it * FZ_THREADS;
                if (e < HJ) {
                    float du = 0.f;
                    if (fnv[it] > 0.f) {
                        du = (dFv[it] - dot * fnv[it]) / nn;
                        s_kap += du * (-__ldcg(data + Y.Fg + (int64_t)c.g * nF + (size_t)e * FZ_K + k) - kaps);
                        s_kaps += du * (-kap);
                    }
                    s.duK[e] = du;
                    dFg_g[(size_t)k * HJ + e] = -kap * du;                 // k-major: a filter's 1200 entries are contiguous for the dots and the tconv below
                }
            }
            s_kap = block_sum512(s_kap, s.red); s_kaps = block_sum512(s_kaps, s.red);
            if (threadIdx.x == 0) { s.dscp[Y.i_kap] += s_kap; s.dscp[Y.i_kaps] += s_kaps; }
        }
        FZ_T(4);
        group_barrier(c);                                         // B2: d Fg is published
        FZ_T(5);
        // D chain of the sequence: d rec = recon(z, y; d G); d z,y += corr_sig(rec + S; d G) + corr_sig(d rec; D); d D += dgrad(z, y; d rec)
        {
            const float* recg = data + SEQ(Y.rec, nS);
            for (int t = threadIdx.x; t < nq4; t += FZ_THREADS) {
                const int q = q0 + (t >> 2);
                s.sigB[t] = __ldcg(recg + 4 * q0 + t) + (s.b[q - c.p0] == (t & 3) ? 1.f : 0.f);
            }
            if (c.nr > 0) recon_rows(c, sv, lo7, hi7, q0, q1, 0.f, nullptr, 0, 0, Lb);
            __syncthreads();
            d_data_step(c, zown, yown, s.sig, s.Dc, s.sigB, s.dG, s.sig, s.Gp, s.dz, s.dy);
        }
        FZ_T(6);
        // F chain of the sequence
        dots_kept(c, data + SEQ(Y.e, nZY), dFg_g, s.kl, cnt, s.dxv, 1, HJ);                // d x += corr2d(e; d Fg)
        FZ_T(7);
        tconv_list(c, sv, dFg_g, s.li2, s.lv2, cnt2, LIST_CAP, nullptr, l, s.u, FZ_M2, 1, HJ);      // u = d e = tconv(x; d Fg)
        {
            float* scr = grad + SEQ(Y.e, nZY) + (size_t)c.p0 * FZ_M2;
            for (int o = threadIdx.x; o < c.nr * FZ_M2; o += FZ_THREADS) {
                const float uv = s.u[o], th_old = s.dth[o];
                float dzy = s.dzyF[o] - uv;
                if (n >= 1) { const float th_new = th_old - uv; s.dth[o] = th_new; dzy -= th_new; }
                s.dzyF[o] = dzy;
                if (has_w) scr[o] = n == 0 ? uv : th_old;
            }
            if (has_w) for (int o = threadIdx.x; o < FZ_NTGT * FZ_M2; o += FZ_THREADS) s.dFt[o] = 0.f;
        }
        FZ_T(8);
        cluster_barrier();                                        // w rows and the dgrad partials of the cluster are complete
        cluster_reduce_Gp();
        if (has_w) {
            dots_kept(c, grad + SEQ(Y.e, nZY), data + Y.F_in + (int64_t)c.g * Y.F_in_gs, s.kl, cnt, s.dxv);      // d x += corr2d(w; F)
            fgrad_targets(c, grad + SEQ(Y.e, nZY), s.li2, s.lv2, cnt2, s.dFt);                               // d F += fgrad(w, x)
            publish_dFt();
        }
        FZ_T(9);
        cluster_barrier();
    }
    // ---- final mask (model.jl:206-210) and the hand-over to the XYZ reverse pass ----------------------------------------------------------
    {
        const float med = __ldcg(data + P.medF + c.g);
        float* gz_ = grad + SEQ(XL.z_out, nZ) + (size_t)c.p0 * FZ_M; float* gy_ = grad + SEQ(XL.y_out, nZ) + (size_t)c.p0 * FZ_M;
        for (int o = threadIdx.x; o < c.nr * FZ_M; o += FZ_THREADS) {
            const int pl = o / FZ_M, m = o - pl * FZ_M;
            const float zv = zown[o], yv = yown[o];
            gz_[o] = s.dz[o] + (zv >= med ? mf * s.dzyF[pl * FZ_M2 + m] : 0.f);
            gy_[o] = s.dy[o] + (yv >= med ? mf * s.dzyF[pl * FZ_M2 + FZ_M + m] : 0.f);
        }
        float* gx = grad + SEQ(XL.x_out, nX);
        for (int q = threadIdx.x; q < cnt; q += FZ_THREADS) if ((q % FZ_CL) == c.r) gx[s.kl[q]] = s.dxv[q];
        if (threadIdx.x < 64) part[nD + threadIdx.x] = s.dscp[threadIdx.x];
    }
    group_barrier(c);
    {
        float* gs = gsum2 + (size_t)c.g * (nF + nD + 64);
        if (c.gidx < FZ_K) {                                      // d F0 column k = the group-level part + the sequences' fgrad(u_0, x)
            const int k = c.gidx;
            for (int e = threadIdx.x; e < HJ; e += FZ_THREADS) {
                const int a = e / FZ_M2, j = e - a * FZ_M2;
                float acc = s.duK[e];
                for (int q = 0; q < d.B; ++q) acc += __ldcg(W.dFp + ((size_t)(c.g * d.B + q) * (FZ_H * FZ_K) + a * FZ_K + k) * FZ_M2 + j);
                gs[(size_t)e * FZ_K + k] = acc;
            }
        }
        if (c.gidx == c.ng - 1) {
            for (int o = threadIdx.x; o < nD; o += FZ_THREADS) {
                float acc = s.dDg[o];
                for (int q = 0; q < d.B; ++q) acc += __ldcg(gpart + (size_t)q * FZ_CL * FZ_PART + o);
                gs[nF + o] = acc;
            }
            if (threadIdx.x < 64) {
                float acc = 0.f;
                for (int q = 0; q < c.ng; ++q) acc += __ldcg(gpart + (size_t)q * FZ_PART + nD + threadIdx.x);
                gs[nF + nD + threadIdx.x] = acc;
            }
        }
    }
#ifdef FZ_PROFILE
    FZ_T(10);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const char* nm[11] = {"prologue", "loss", "B1", "d_update adj", "f_update adj", "B2", "D chain", "dots(e,dFg)", "tconv(dFg)+elem", "bar+dots(w)+fgrad", "epilogue"};
        long long tot = 0; for (int i = 0; i < 11; ++i) tot += fz_acc[i];
        for (int i = 0; i < 11; ++i) printf("[fzd] %-20s %9lld clk %5.1f%%\n", nm[i], fz_acc[i], 100.0 * fz_acc[i] / tot);
        printf("[fzd] total %lld clk\n", tot);
    }
#endif
#undef SEQ
}

// adds the group sums of the fused reverse pass to the tape's adjoints of the prepared filters / scalars (deterministic: groups in order)
__global__ void __launch_bounds__(256) k_csc_fused_finish(const float* __restrict__ gsum, int G, int nF, int nD, int nsc, float* __restrict__ gF, float* __restrict__ gD, float* __restrict__ gsc) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int tot = nF + nD + 64;
    if (o >= tot) return;
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc += gsum[(size_t)g * tot + o];
    if (o < nF) gF[o] += acc;
    else if (o < nF + nD) gD[o - nF] += acc;
    else if (o - nF - nD < nsc) gsc[o - nF - nD] += acc;
}

// Last kernel of the fully fused reverse pass: group sums of the two reverse kernels (gsum: [G2][nF + nD + 64]) pushed through the adjoint of
// prep_params (model.jl:139-169: F = r^2 / ||r^2||_2 per syntax filter, D = (r^2 + 1e-3) normalised over the four nucleotides, scalars
// squared) into the raw gradient vector (assigned: the step has no other contributor).  Blocks 0..K-1: one syntax filter each; block K: D;
// block K+1: the trained scalars; the remaining blocks clear the sync area (barrier counters, median histograms) for the next step.
// Replaces k_csc_fused_finish, the three prep adjoint kernels of the tape and four memset nodes.
__global__ void __launch_bounds__(256) k_csc_fused_tail(const float* __restrict__ gsum, int G2, const float* __restrict__ raw, int64_t off_D, int64_t off_F,
                                                        const float* __restrict__ Feff, const float* __restrict__ Fnrm, const float* __restrict__ Deff,
                                                        float* __restrict__ draw, ScalarSegs sg, uint4* __restrict__ zero_area, int64_t zero_n16, CscDims d) {
    __shared__ float s_de[FZ_H * FZ_M2];
    __shared__ float s_dot;
    const int nF = d.h * d.M2 * d.K, nD = d.f_len * d.M, tot = nF + nD + 64, HJ = d.h * d.M2;
    if ((int)blockIdx.x < d.K) {
        const int k = blockIdx.x;
        float dot = 0.f;
        for (int e = threadIdx.x; e < HJ; e += blockDim.x) {           // e = j*h + a in raw order
            const int a = e % d.h, j = e / d.h;
            const int64_t o = ((int64_t)a * d.M2 + j) * d.K + k;
            float de = 0.f;
            for (int g = 0; g < G2; ++g) de += gsum[(size_t)g * tot + o];
            s_de[e] = de; dot += de * Feff[o];
        }
        dot = block_sum(dot);
        if (threadIdx.x == 0) s_dot = dot;
        __syncthreads();
        dot = s_dot;
        const float nn = Fnrm[k];
        for (int e = threadIdx.x; e < HJ; e += blockDim.x) {
            const int a = e % d.h, j = e / d.h;
            const int64_t o = ((int64_t)a * d.M2 + j) * d.K + k;
            const float r = raw[off_F + (int64_t)k * HJ + e];
            draw[off_F + (int64_t)k * HJ + e] = (s_de[e] - dot * Feff[o]) / nn * 2.f * r;
        }
    } else if ((int)blockIdx.x == d.K) {
        for (int t = threadIdx.x; t < d.fl * d.M; t += blockDim.x) {
            const int m = t % d.M, j = t / d.M;
            float de[4], ssum = 0.f, dot = 0.f;
            #pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int o = (4 * j + a) * d.M + m;
                const float r = raw[off_D + m * d.f_len + 4 * j + a];
                ssum += r * r + 0.001f;
                float v = 0.f;
                for (int g = 0; g < G2; ++g) v += gsum[(size_t)g * tot + nF + o];
                de[a] = v; dot += v * Deff[o];
            }
            #pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float r = raw[off_D + m * d.f_len + 4 * j + a];
                draw[off_D + m * d.f_len + 4 * j + a] = (de[a] - dot) / ssum * 2.f * r;
            }
        }
    } else if ((int)blockIdx.x == d.K + 1) {
        const int w = threadIdx.x >> 5;
        if (w < sg.nseg)
            for (int i = threadIdx.x & 31; i < sg.n[w]; i += 32) {
                const int idx = sg.eff_idx[w] + i;
                float de = 0.f;
                for (int g = 0; g < G2; ++g) de += gsum[(size_t)g * tot + nF + nD + idx];
                draw[sg.raw_off[w] + i] = 2.f * raw[sg.raw_off[w] + i] * de;
            }
    } else {
        const int64_t nb = (int64_t)gridDim.x - (d.K + 2);
        for (int64_t i = ((int64_t)blockIdx.x - (d.K + 2)) * blockDim.x + threadIdx.x; i < zero_n16; i += nb * blockDim.x) zero_area[i] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// First kernel of the fused step: prep_params (model.jl:139-169) of the three parameter families and the 2-bit -> byte unpacking of the batch's
// bases in one launch (blocks 0..K-1: syntax filters; K: D; K+1: scalars; the rest: bases).  Replaces four kernels of the tape.
__global__ void __launch_bounds__(256) k_csc_fused_head(const float* __restrict__ raw, int64_t off_D, int64_t off_F, float* __restrict__ Feff, float* __restrict__ Fnrm,
                                                        float* __restrict__ Deff, float* __restrict__ sceff, ScalarSegs sg,
                                                        const uint32_t* __restrict__ words, int64_t rowwords, const int64_t* __restrict__ idx, uint8_t* __restrict__ bases, CscDims d) {
    __shared__ float s_n;
    const int HJ = d.h * d.M2;
    if ((int)blockIdx.x < d.K) {
        const int k = blockIdx.x;
        float ss = 0.f;
        for (int e = threadIdx.x; e < HJ; e += blockDim.x) { const float r = raw[off_F + (int64_t)k * HJ + e]; const float u = r * r; ss += u * u; }
        ss = block_sum(ss);
        if (threadIdx.x == 0) { s_n = sqrtf(ss); Fnrm[k] = s_n; }
        __syncthreads();
        for (int e = threadIdx.x; e < HJ; e += blockDim.x) {
            const int a = e % d.h, j = e / d.h;
            const float r = raw[off_F + (int64_t)k * HJ + e];
            Feff[((int64_t)a * d.M2 + j) * d.K + k] = r * r / s_n;
        }
    } else if ((int)blockIdx.x == d.K) {
        for (int t = threadIdx.x; t < d.fl * d.M; t += blockDim.x) {
            const int m = t % d.M, j = t / d.M;
            float u[4], ssum = 0.f;
            #pragma unroll
            for (int a = 0; a < 4; ++a) { const float r = raw[off_D + m * d.f_len + 4 * j + a]; u[a] = r * r + 0.001f; ssum += u[a]; }
            #pragma unroll
            for (int a = 0; a < 4; ++a) Deff[(4 * j + a) * d.M + m] = u[a] / ssum;
        }
    } else if ((int)blockIdx.x == d.K + 1) {
        const int w = threadIdx.x >> 5;
        if (w < sg.nseg)
            for (int i = threadIdx.x & 31; i < sg.n[w]; i += 32) { const float r = raw[sg.raw_off[w] + i]; sceff[sg.eff_idx[w] + i] = r * r; }
    } else {
        const int64_t nb = (int64_t)gridDim.x - (d.K + 2), tot = (int64_t)d.NS * d.Lb;
        for (int64_t t = ((int64_t)blockIdx.x - (d.K + 2)) * blockDim.x + threadIdx.x; t < tot; t += nb * blockDim.x) {
            const int64_t n = t / d.Lb; const int p = (int)(t - n * d.Lb);
            bases[t] = (uint8_t)((words[idx[n] * rowwords + (p >> 4)] >> (2 * (p & 15))) & 3u);
        }
    }
}
