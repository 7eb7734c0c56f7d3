// CTA-pair version of k_scan_tc (scan_tc.cuh): tcgen05.mma.cta_group::2, M = 256.
//
// Why: in k_scan_tc every 128x256x16 MMA fetches 4 KB of A and 8 KB of B from the SM's shared memory for 128 clocks of math; ncu shows
// the operand path (sm__mem_tensor_cycles_active 78 %) pacing the tensor pipe together with the accumulator drains, and every change
// that raised the bytes per flop (N = 128 MMAs) made the kernel slower.  A CTA pair halves the B traffic: the two CTAs of a cluster
// work on two different position tiles against the SAME 256 slots, each stages only half of B (128 slots) and the pair's MMA
// (issued by the leader CTA) multiplies both A tiles with the whole B: 8 KB instead of 12 KB per MMA and SM.
//
// Differences to k_scan_tc, everything else (operand layout, epilogue, candidate list) is the same code:
//   * cluster of 2 CTAs; rank r works on tile 2*dt + r of the pair's double tile dt; only rank 0 has an active MMA warp
//   * leader's full[stage] barrier counts both producers (the peer arrives remotely, release.cluster / acquire.cluster)
//   * tcgen05.commit multicasts to the barriers of both CTAs (stage free, accumulator full)
//   * the leader's accumulator-empty barriers count the epilogue warps of both CTAs (32 arrivals, the peer's are remote)
//   * TMEM is allocated / released with cta_group::2, cluster barriers around set-up and tear-down
#pragma once

__device__ __forceinline__ uint32_t tcs_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tcs_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t tcs_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void tcs_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tcs_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// wait with cluster-scope acquire (the arrivals may come from the peer CTA)
__device__ __forceinline__ void tcs_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    } while (!done);
}
__device__ __forceinline__ void tcs_mma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xFFFFFFFF;\n\t@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all MMAs issued so far -> the same barrier in both CTAs of the pair
__device__ __forceinline__ void tcs_commit2(uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\telect.sync _|q, 0xFFFFFFFF;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" :: "r"(bar) : "memory");
}
template <int KP>
__device__ __forceinline__ void tcs_issue2_n(uint32_t tmem_d, uint64_t da0, uint64_t db0, uint32_t idesc) {
    #pragma unroll
    for (int t = 0; t < KP; ++t) tcs_mma2(tmem_d, da0 + (uint64_t)(2 * t), db0 + (uint64_t)(2 * t * (TCS_N / 2)), idesc, t ? 1u : 0u);
}
__device__ __forceinline__ void tcs_issue2(int kp, uint32_t tmem_d, uint64_t da0, uint64_t db0, uint32_t idesc) {
    switch (kp) {
        case 1: tcs_issue2_n<1>(tmem_d, da0, db0, idesc); break;   case 2: tcs_issue2_n<2>(tmem_d, da0, db0, idesc); break;
        case 3: tcs_issue2_n<3>(tmem_d, da0, db0, idesc); break;   case 4: tcs_issue2_n<4>(tmem_d, da0, db0, idesc); break;
        case 5: tcs_issue2_n<5>(tmem_d, da0, db0, idesc); break;   case 6: tcs_issue2_n<6>(tmem_d, da0, db0, idesc); break;
        case 7: tcs_issue2_n<7>(tmem_d, da0, db0, idesc); break;   case 8: tcs_issue2_n<8>(tmem_d, da0, db0, idesc); break;
        case 9: tcs_issue2_n<9>(tmem_d, da0, db0, idesc); break;   case 10: tcs_issue2_n<10>(tmem_d, da0, db0, idesc); break;
        case 11: tcs_issue2_n<11>(tmem_d, da0, db0, idesc); break; case 12: tcs_issue2_n<12>(tmem_d, da0, db0, idesc); break;
        case 13: tcs_issue2_n<13>(tmem_d, da0, db0, idesc); break; case 14: tcs_issue2_n<14>(tmem_d, da0, db0, idesc); break;
        case 15: tcs_issue2_n<15>(tmem_d, da0, db0, idesc); break; default: tcs_issue2_n<16>(tmem_d, da0, db0, idesc); break;
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TCS_THREADS, 1) k_scan_tc2(const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t tcs_smem[];
    uint8_t* sA = tcs_smem;                                              // [TCS_STAGES][2 parities][2 planes][TCS_STREAM][16 B]
    uint8_t* sB = tcs_smem + TCS_STAGES * TCS_STAGE_BYTES;               // per sub-block [kchunks][128 slots of this CTA][16 B]
    __shared__ __align__(8) uint64_t s_bars[2 * TCS_STAGES + 5];         // full[S], empty[S], accfull[2], accempty[2], B landed
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = tcs_cluster_rank();                           // 0 = leader
    auto bar = [&](int i) { return (uint32_t)__cvta_generic_to_shared(&s_bars[i]); };

    int bi = 0;
    for (int i = 0; i < a.nblocks; ++i) if ((int)blockIdx.x >= a.blocks[i].cta0) bi = i;
    const TcBlock blk = a.blocks[bi];
    const int prank = ((int)blockIdx.x - blk.cta0) >> 1, npairs = blk.nctas >> 1;      // pair index inside the entry
    const int ndt = (a.ntiles + 1) >> 1;                                               // double tiles

    if (tid == 0) {
        for (int i = 0; i < TCS_STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(bar(i)) : "memory");                       // both producers
        for (int i = TCS_STAGES; i < 2 * TCS_STAGES + 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar(i)) : "memory");      // commits
        for (int i = 2 * TCS_STAGES + 2; i < 2 * TCS_STAGES + 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" :: "r"(bar(i)) : "memory");   // epilogue warps of both CTAs
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar(2 * TCS_STAGES + 4)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < TCS_STAGES * 2; i += blockDim.x)
        *reinterpret_cast<uint4*>(sA + (size_t)i * (2 * TCS_PLANE_BYTES) + TCS_PLANE_BYTES + (TCS_STREAM - 1) * 16) = make_uint4(0, 0, 0, 0);
    const uint32_t sA_addr = (uint32_t)__cvta_generic_to_shared(sA), sB_addr = (uint32_t)__cvta_generic_to_shared(sB);
    const uint32_t b_bytes0 = (uint32_t)blk.kchunks[0] * (TCS_N / 2) * 16;
    __syncthreads();                                                     // barrier inits visible to this CTA's threads
    if (warp == 1 && lane == 0) {
        // this CTA's half of the B operands: slots [128 crank, 128 crank + 128) of every K chunk (2 KB pieces)
        const uint32_t b_bytes1 = blk.nsub > 1 ? (uint32_t)blk.kchunks[1] * (TCS_N / 2) * 16 : 0u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar(2 * TCS_STAGES + 4)), "r"(b_bytes0 + b_bytes1) : "memory");
        for (int sub = 0; sub < blk.nsub; ++sub)
            for (int c = 0; c < blk.kchunks[sub]; ++c)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(sB_addr + (sub ? b_bytes0 : 0u) + (uint32_t)c * (TCS_N / 2) * 16),
                                "l"(a.blob + blk.b_off[sub] + ((size_t)c * TCS_N + (size_t)crank * (TCS_N / 2)) * 16), "r"((uint32_t)(TCS_N / 2) * 16),
                                "r"(bar(2 * TCS_STAGES + 4)) : "memory");
    }
    tcs_wait(bar(2 * TCS_STAGES + 4), 0);                                // every thread: this CTA's B has landed
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tcs_cluster_sync();                                                  // both CTAs: barriers initialised, TMEM allocated, B resident
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    // instruction descriptor: D = F32, A = B = F16, K-major, N = 256, M = 256 (the pair)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(TCS_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

    if (warp == 0 || warp == TCS_THREADS / 32 - 1) {
        // ---- producers (as in k_scan_tc); the full barrier lives in the leader CTA ----
        const int pw = warp == 0 ? 0 : 1;
        int it = 0;
        for (int dt = prank; dt < ndt; dt += npairs, ++it) {
            if ((it & 1) != pw) continue;
            const int tile = 2 * dt + (int)crank;
            const int st = it % TCS_STAGES; const uint32_t ph = (it / TCS_STAGES) & 1;
            uint32_t v = (uint32_t)tile * 256u + 2u * lane;
            uint32_t n = v / a.Lb, p = v - n * a.Lb;
            const uint32_t* row = a.seqw + (a.seq0 + n) * a.rowwords;
            uint32_t b0[TCS_STREAM / 32], b1[TCS_STREAM / 32];
            #pragma unroll
            for (int i = 0; i < TCS_STREAM / 32; ++i) {
                uint32_t w0 = 0, w1 = 0;
                const bool second_in_row = p + 1 < a.Lb;
                const uint32_t p1 = second_in_row ? p + 1 : 0u;
                const uint32_t* row1 = second_in_row ? row : row + a.rowwords;
                if (v < a.vtotal) w0 = __ldg(row + (p >> 4));
                if (v + 1 < a.vtotal) w1 = __ldg(row1 + (p1 >> 4));
                b0[i] = (w0 >> ((p & 15) * 2)) & 3u;
                b1[i] = (w1 >> ((p1 & 15) * 2)) & 3u;
                v += 64; p += 64;
                while (p >= a.Lb) { p -= a.Lb; row += a.rowwords; }
            }
            tcs_wait(bar(TCS_STAGES + st), ph ^ 1);                                      // stage free (commit multicast reaches both CTAs)
            uint8_t* e0 = sA + (size_t)st * TCS_STAGE_BYTES;
            uint8_t* o0 = e0 + 2 * TCS_PLANE_BYTES;
            #pragma unroll
            for (int i = 0; i < TCS_STREAM / 32; ++i) {
                const uint32_t m = lane + 32 * i;
                uint32_t b2 = __shfl_down_sync(0xffffffffu, b0[i], 1);
                const uint32_t nxt = __shfl_sync(0xffffffffu, i + 1 < TCS_STREAM / 32 ? b0[(i + 1) % (TCS_STREAM / 32)] : 0u, 0);
                if (lane == 31) b2 = nxt;
                const uint4 ev = tcs_onehot(b0[i], b1[i]);
                const uint4 od = tcs_onehot(b1[i], b2);
                *reinterpret_cast<uint4*>(e0 + m * 16) = ev;
                *reinterpret_cast<uint4*>(o0 + m * 16) = od;
                if (m > 0) {
                    *reinterpret_cast<uint4*>(e0 + TCS_PLANE_BYTES + (m - 1) * 16) = ev;
                    *reinterpret_cast<uint4*>(o0 + TCS_PLANE_BYTES + (m - 1) * 16) = od;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (crank == 0) tcs_arrive_local(bar(st));
                else tcs_arrive_remote(tcs_mapa(bar(st), 0));
            }
        }
    } else if (warp == 1) {
        if (crank == 0) {   // ---- MMA warp of the leader: converged, tcs_mma2 / tcs_commit2 elect the issuing lane ----
            const uint64_t dbA = tcs_desc(sB_addr, (TCS_N / 2) * 16, 128), dbB = tcs_desc(sB_addr + b_bytes0, (TCS_N / 2) * 16, 128);
            const int kpA = blk.kchunks[0] >> 1, kpB = blk.kchunks[1] >> 1;
            const long long t_start = clock64();
            uint32_t u = 0;
            int it = 0;
            for (int dt = prank; dt < ndt; dt += npairs, ++it) {
                const int st = it % TCS_STAGES; const uint32_t ph = (it / TCS_STAGES) & 1;
                tcs_wait_cluster(bar(st), ph);                                           // both CTAs' streams built
                #pragma unroll
                for (int par = 0; par < 2; ++par)
                    for (int sub = 0; sub < blk.nsub; ++sub, ++u) {
                        const uint32_t ac = u & 1u, uses = u >> 1;
                        tcs_wait_cluster(bar(2 * TCS_STAGES + 2 + ac), (uses & 1) ^ 1);  // drained in both CTAs
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t da0 = tcs_desc(sA_addr + st * TCS_STAGE_BYTES + par * 2 * TCS_PLANE_BYTES, TCS_PLANE_BYTES, 128);
                        tcs_issue2(sub ? kpB : kpA, tmem + ac * TCS_N, da0, sub ? dbB : dbA, idesc);
                        tcs_commit2(bar(2 * TCS_STAGES + ac));
                    }
                tcs_commit2(bar(TCS_STAGES + st));
            }
            if (a.clocks && lane == 0) { a.clocks[blockIdx.x * 2] = clock64() - t_start; a.clocks[blockIdx.x * 2 + 1] = 2 * it; }
        } else if (a.clocks && lane == 0) { a.clocks[blockIdx.x * 2] = 0; a.clocks[blockIdx.x * 2 + 1] = 0; }
    } else {
        // ---- epilogue: as in k_scan_tc, on this CTA's own tile and TMEM; the accumulator-empty barrier lives in the leader ----
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;
        unsigned long long cur_end[2] = {0, 0};
        uint32_t dead = 0;
        uint32_t u = 0;
        for (int dt = prank; dt < ndt; dt += npairs)
        for (int par = 0; par < 2; ++par)
        for (int sub = 0; sub < blk.nsub; ++sub, ++u) {
            const uint32_t ac = u & 1u, uses = u >> 1;
            const uint32_t v = (uint32_t)(2 * dt + (int)crank) * 256u + (uint32_t)par + 2u * (uint32_t)(q * 32 + lane);
            const bool inb = v < a.vtotal;
            tcs_wait(bar(2 * TCS_STAGES + ac), uses & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem + ac * TCS_N + cg * 64 + ((uint32_t)(q * 32) << 16);
            uint32_t ua[32], ub[32];
            TCS_LDTM32(ua, taddr);
            TCS_LDTM32(ub, taddr + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (crank == 0) tcs_arrive_local(bar(2 * TCS_STAGES + 2 + ac));
                else tcs_arrive_remote(tcs_mapa(bar(2 * TCS_STAGES + 2 + ac), 0));
            }
            const float ma = tcs_max32(ua), mb = tcs_max32(ub);
            const bool pos = inb && fmaxf(ma, mb) > 0.f;
            if (__any_sync(0xffffffffu, pos)) {
                uint32_t cb0 = 0u, cb1 = 0u;
                if (__any_sync(0xffffffffu, inb && ma > 0.f)) cb0 = pos ? tcs_posbits(ua) : 0u;
                if (__any_sync(0xffffffffu, inb && mb > 0.f)) cb1 = pos ? tcs_posbits(ub) : 0u;
                tcs_append(a.list, a.gcount, a.cap, a.overflow, lane, cb0, cb1, (uint32_t)(blk.slot0[sub] + cg * 64), v, cur_end, &dead);
            }
        }
        for (unsigned long long i = cur_end[0] + lane; i < cur_end[1]; i += 32) a.list[i] = ~0ull;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tcs_cluster_sync();                                                  // both CTAs are done with both TMEMs and with remote barriers
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
}
