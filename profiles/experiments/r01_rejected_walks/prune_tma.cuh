// prune_tma.cuh -- the Q-specialised, variant-only pruning walk with its operands staged by TMA bulk copies.
//
// Same arithmetic, op list and results (bit for bit) as k_prune<false, true, true>.  The plain walk is latency bound and
// a register pipeline cannot be expressed (ptxas puts every global load on one scoreboard, prune_rp.cuh).  Here the
// operands of the next ops travel by cp.async.bulk (TMA, 1-D): a warp owns 16 consecutive sites, so an internal child is
// ONE contiguous 2 KB copy (+ 128 B of scales) and a leaf child one 512 B copy, issued by one lane into a warp-private
// ring of STAGES shared-memory stages, each with its own mbarrier (expect_tx / complete_tx).  Completion is tracked per
// stage, not per scoreboard: up to STAGES - 1 ops of loads are in flight per warp while it computes, the copies never
// touch the LSU/L1 path of the lanes, and the lanes fetch their 32-byte slices with two conflict-free LDS.128.
//
// Hazards.  An operand that an earlier op of the same launch writes is NOT staged: the host flags it (SV_OP_DEP1 /
// SV_OP_DEP2) and the lanes load it themselves when the op is computed, exactly like k_prune (a lane only re-reads what
// it wrote itself, so no fence is needed).  Everything else -- the leaves, and partials left by earlier launches -- is
// read-only for the launch and can be requested any number of ops ahead.  (The first version staged the dependent
// operands too, ordered by a host-computed `ready_after` and published with fence.proxy.async after the producing op:
// that fence compiles to MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC and took 12 % of all stall samples.)  The forwarded first
// child, the common dependency of a post-order walk, stays in registers as in k_prune.
//
// Needs an even site capacity (16-byte alignment of the scale rows); the host falls back to k_prune otherwise.
//
// MEASURED (B200, 1 000 cells x 100 000 sites, SIEVE_B200_TMA=2 / 3 / 4 stages), bit-identical every time
// (tests/test_gpu_parity.py::test_pipelined_walk_is_bit_identical, resident and streamed), against 2.50 ms for the plain walk:
//   first version (dependent operands staged too, fence per producing op, half-swapped slices): 5.7 / 5.6 / 7.3 ms.
//     ncu (profiles/r01k_walk_tma_ncu.txt): the pipeline works -- long-scoreboard stalls per issue drop from 4.8 to 1.6 --
//     but 2.69e9 instructions against 1.62e9 (431 against 260 per op and warp: every lane decodes the descriptors it
//     requests, 64-bit address arithmetic, the selects of the slices) and the fence on a third of the ops;
//   this version: 4.2 / 4.9 / 6.2 ms.  Still SLOWER, so opt-in; and more stages are slower, i.e. the occupancy that the
//     stages cost (52 KB per block at 3 stages: 4 blocks/SM, 94 registers) is worth more than the look-ahead they buy.
// What is left to try: a producer warp per block instead of request logic in every warp, operands packed so that one is
// one copy (scales next to the vectors), 8 sites per warp and stage to halve the shared memory.
#pragma once
#include "prune.cuh"

#define SV_TMA_STAGE_BYTES 4352  // [child 1: 16 sites x 128 B][child 2: 2048 B][scales of child 1: 128 B][of child 2: 128 B]
#define SV_TMA_WARPS (SV_PRUNE_THREADS / 32)
__host__ __device__ constexpr int sv_tma_smem(int stages) { return SV_TMA_WARPS * stages * (SV_TMA_STAGE_BYTES + 8); }

__device__ __forceinline__ void sv_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sv_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sv_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool sv_mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void sv_lds128(unsigned addr, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
// a lane's 32-byte slice as two 16-byte halves (2-way bank conflicts inside a quarter warp: the two sites of a quarter
// warp are 128 bytes apart; swapping the halves by site parity avoids them but costs eight selects per slice)
__device__ __forceinline__ void sv_lds_slice(unsigned addr, double (&v)[4]) {
    sv_lds128(addr, v[0], v[1]);
    sv_lds128(addr + 16, v[2], v[3]);
}
__device__ __forceinline__ double sv_lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

template <bool GENO0, int STAGES>
__global__ void __launch_bounds__(SV_PRUNE_THREADS) k_prune_tma(SvPruneArgs a) {
    constexpr int NS = 2;
    extern __shared__ __align__(128) unsigned char sv_tma_smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int k = min(lane & 3, a.K - 1);
    const int klane = lane & 3;
    const int site_w = blockIdx.x * (SV_TMA_WARPS * 16) + warp * 16;  // first site of this warp
    if (site_w >= a.S) return;                                         // no block-wide barrier anywhere
    const int n_valid = min(16, a.S - site_w);                          // even: the host launches this kernel for even S only
    const int site0 = site_w + (lane >> 2);
    bool site_ok[NS];
    unsigned srow[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const int site = site0 + 8 * j;
        site_ok[j] = site < a.S;
        srow[j] = (unsigned)(site_ok[j] ? site : a.S - 1);
    }
    const int rg = GENO0 ? 0 : a.root_genotype;
    const int i0 = a.op_begin, i1 = a.op_end;
    if (i0 >= i1) return;

    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(sv_tma_smem_raw);
    const unsigned stage0 = smem0 + (unsigned)warp * (STAGES * SV_TMA_STAGE_BYTES);
    const unsigned bar0 = smem0 + SV_TMA_WARPS * STAGES * SV_TMA_STAGE_BYTES + (unsigned)warp * (STAGES * 8);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) sv_mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncwarp();

    // this lane's slices inside a stage: site j of the lane, category k -> 32 bytes
    unsigned voff[NS], loff[NS], soff[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const unsigned sl = (unsigned)(lane >> 2) + 8 * j;
        voff[j] = sl * 128 + (unsigned)k * 32;  // internal child: [site][K][4]
        loff[j] = sl * 32;                      // leaf child: [site][4]
        soff[j] = sl * 8;
    }

    // requests the memory operands of op p into its stage (every lane evaluates the descriptor, lane 0 talks to the TMA)
    auto issue = [&](const int p) {
        const int4* opp = reinterpret_cast<const int4*>(a.ops + p);
        const int4 d0 = __ldg(opp), d1 = __ldg(opp + 1);
        const int flags = d1.y;
        const unsigned st = (unsigned)((p - i0) % STAGES);
        const unsigned base = stage0 + st * SV_TMA_STAGE_BYTES;
        const unsigned bar = bar0 + 8 * st;
        const bool need1 = !(flags & (SV_OP_FWD1 | SV_OP_DEP1)), need2 = !(flags & (SV_OP_UNARY | SV_OP_DEP2));
        const bool leaf1 = flags & SV_OP_LEAF1, leaf2 = flags & SV_OP_LEAF2;
        unsigned bytes = 0;
        if (need1) bytes += leaf1 ? (unsigned)n_valid * 32 : (unsigned)n_valid * 136;
        if (need2) bytes += leaf2 ? (unsigned)n_valid * 32 : (unsigned)n_valid * 136;
        if (lane == 0) {
            sv_mbar_expect_tx(bar, bytes);
            if (need1) {
                const size_t row = (size_t)(unsigned)d0.y + (unsigned)site_w;
                if (leaf1) {
                    sv_bulk_g2s(base, a.leafX + row * 4, (unsigned)n_valid * 32, bar);
                } else {
                    sv_bulk_g2s(base, a.part + row * 16, (unsigned)n_valid * 128, bar);
                    sv_bulk_g2s(base + 4096, a.scale + row, (unsigned)n_valid * 8, bar);
                }
            }
            if (need2) {
                const size_t row = (size_t)(unsigned)d0.z + (unsigned)site_w;
                if (leaf2) {
                    sv_bulk_g2s(base + 2048, a.leafX + row * 4, (unsigned)n_valid * 32, bar);
                } else {
                    sv_bulk_g2s(base + 2048, a.part + row * 16, (unsigned)n_valid * 128, bar);
                    sv_bulk_g2s(base + 4096 + 128, a.scale + row, (unsigned)n_valid * 8, bar);
                }
            }
        }
    };

    double x[NS][4], xs[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        xs[j] = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) x[j][g] = 0.0;
    }
    for (int p = i0; p < i1 && p < i0 + STAGES - 1; p++) issue(p);  // the first STAGES - 1 ops are requested up front
    int4 n0 = __ldg(reinterpret_cast<const int4*>(a.ops + i0));
    int4 n1 = __ldg(reinterpret_cast<const int4*>(a.ops + i0) + 1);
    double2 qn1 = __ldg(a.qfac + (size_t)n0.w * SV_KMAX + k);
    double2 qn2 = (n1.y & SV_OP_UNARY) ? make_double2(0.0, 0.0) : __ldg(a.qfac + (size_t)n1.x * SV_KMAX + k);

    for (int i = i0; i < i1; i++) {
        // the stage of op i - 1 was released by the __syncwarp of its op: it takes the request for op i + STAGES - 1
        if (i + STAGES - 1 < i1) issue(i + STAGES - 1);
        const int4 o0 = n0, o1 = n1;
        const double2 q1 = qn1, q2 = qn2;
        if (i + 1 < i1) {
            n0 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1));
            n1 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1) + 1);
            qn1 = __ldg(a.qfac + (size_t)n0.w * SV_KMAX + k);
            qn2 = (n1.y & SV_OP_UNARY) ? make_double2(0.0, 0.0) : __ldg(a.qfac + (size_t)n1.x * SV_KMAX + k);
        }
        const unsigned dst_base = (unsigned)o0.x;
        const int flags = o1.y;
        const bool unary = flags & SV_OP_UNARY;
        const unsigned st = (unsigned)((i - i0) % STAGES);
        const unsigned base = stage0 + st * SV_TMA_STAGE_BYTES;
        const unsigned parity = (unsigned)(((i - i0) / STAGES) & 1);
        {
            unsigned spins = 0;
            while (!sv_mbar_try_wait(bar0 + 8 * st, parity)) {
                if (++spins > (1u << 24)) __trap();  // never hang the device on a protocol error
            }
        }
        double w[NS][4], ws[NS];
#pragma unroll
        for (int j = 0; j < NS; j++) {
            ws[j] = 0.0;
#pragma unroll
            for (int g = 0; g < 4; g++) w[j][g] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < NS; j++) {
            if (!(flags & SV_OP_FWD1)) {
                if (flags & SV_OP_DEP1) {  // written by an earlier op of this launch: the lane's own load, as in k_prune
                    const size_t row = (size_t)(unsigned)o0.y + srow[j];
                    sv_ld256(a.part + row * 16 + k * 4, x[j]);
                    xs[j] = a.scale[row];
                } else if (flags & SV_OP_LEAF1) {
                    sv_lds_slice(base + loff[j], x[j]);
                    xs[j] = 0.0;
                } else {
                    sv_lds_slice(base + voff[j], x[j]);
                    xs[j] = sv_lds64(base + 4096 + soff[j]);
                }
            }
            if (!unary) {
                if (flags & SV_OP_DEP2) {
                    const size_t row = (size_t)(unsigned)o0.z + srow[j];
                    sv_ld256(a.part + row * 16 + k * 4, w[j]);
                    ws[j] = a.scale[row];
                } else if (flags & SV_OP_LEAF2) {
                    sv_lds_slice(base + 2048 + loff[j], w[j]);
                } else {
                    sv_lds_slice(base + 2048 + voff[j], w[j]);
                    ws[j] = sv_lds64(base + 4096 + 128 + soff[j]);
                }
            }
        }
        __syncwarp();  // every lane has its slices: the stage may be overwritten by the next request for it

        const bool cz0 = flags & SV_OP_ZEROD;
        double y[NS][4];
        sv_matvec_q<NS>(q1, x, y, cz0);
        if (!unary) {
            double z[NS][4];
            sv_matvec_q<NS>(q2, w, z, cz0);
#pragma unroll
            for (int j = 0; j < NS; j++)
#pragma unroll
                for (int p = 0; p < 4; p++) y[j][p] *= z[j][p];
        }
#pragma unroll
        for (int j = 0; j < NS; j++) {
            int hi = sv_maxhi4(y[j]);
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
            double sh;
            const double f = sv_pow2_rescale_hi(hi, &sh);
#pragma unroll
            for (int p = 0; p < 4; p++) x[j][p] = y[j][p] * f;
            xs[j] = (xs[j] + ws[j]) + sh;
        }
#pragma unroll
        for (int j = 0; j < NS; j++) {
            if ((flags & SV_OP_NOSTORE) || !site_ok[j]) continue;
            const size_t dst_row = dst_base + srow[j];
            sv_st256(a.part + dst_row * 16 + k * 4, x[j]);
            a.scale[dst_row] = xs[j];
        }
        if (flags & SV_OP_ROOT) {
#pragma unroll
            for (int j = 0; j < NS; j++) {
                double v = (klane < a.K ? a.prop[klane] : 0.0) * (GENO0 ? x[j][0] : sv_sel(x[j], rg));
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                const double lsum = site_ok[j] ? __ldg(a.lsum + a.out_offset + site0 + 8 * j) : 0.0;
                if (site_ok[j] && klane == 0) a.site_logl[a.out_offset + site0 + 8 * j] = log(v) + (xs[j] + lsum);
            }
        }
    }
}
