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
// Hazards.  An operand that an op of the same launch writes must not be requested before that op has stored it: the
// host annotates every op with `ready_after`, the index of the last earlier op that writes one of its memory operands
// (-1: none), and flags that producer SV_OP_FENCE: after its stores the warp issues fence.proxy.async + __syncwarp so
// that the copy engine (async proxy) sees them.  Ops are issued in order; an op that is not ready holds the ones behind
// it back (the forwarded first child, the common dependency of a post-order walk, stays in registers as in k_prune).
//
// Needs an even site capacity (16-byte alignment of the scale rows); the host falls back to k_prune otherwise.
//
// MEASURED (B200, 1 000 cells x 100 000 sites, SIEVE_B200_TMA=2 / 3 / 4 stages): 5.7 / 5.6 / 7.3 ms against 2.50 ms for
// the plain walk, bit-identical (tests/test_gpu_parity.py::test_pipelined_walk_is_bit_identical, resident and streamed).
// SLOWER, so opt-in -- but for a different reason than the earlier pipelines (profiles/r01k_walk_tma_ncu.txt): the
// pipeline works, long-scoreboard stalls per issue drop from 4.8 to 1.6 and DRAM is no longer waited for; what it costs
// is instructions, 2.69e9 against 1.62e9 (every lane decodes the descriptors it requests, 64-bit address arithmetic,
// the half-swapping selects of the slices), and the publishing fence: fence.proxy.async compiles to MEMBAR.ALL.CTA +
// FENCE.VIEW.ASYNC and is executed by a third of the ops (12 % of all stall samples on that one instruction, 11 % more
// on the first instruction after the mbarrier wait).  The version to try next: the request logic under a lane-0 branch
// (or a dedicated producer warp per block), one fence per dependent REQUEST instead of per producing op, scales
// packed next to the vectors so that an operand is one copy.
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
// a lane's 32-byte slice as two 16-byte halves, the half at `first` (0 or 16) fetched first
__device__ __forceinline__ void sv_lds_slice(unsigned addr, unsigned first, double (&v)[4]) {
    double a0, a1, b0, b1;
    sv_lds128(addr + first, a0, a1);
    sv_lds128(addr + (16 - first), b0, b1);
    v[0] = first ? b0 : a0;
    v[1] = first ? b1 : a1;
    v[2] = first ? a0 : b0;
    v[3] = first ? a1 : b1;
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

    // this lane's slices inside a stage: site j of the lane, category k -> 32 bytes, fetched as two 16-byte halves in an
    // order that alternates with the site's parity (no bank conflicts inside a quarter warp)
    const unsigned half0 = ((lane >> 2) & 1) * 16;
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
        const bool need1 = !(flags & SV_OP_FWD1), need2 = !(flags & SV_OP_UNARY);
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
    auto ready_after = [&](const int p) { return __ldg(&a.ops[p].pad0); };

    double x[NS][4], xs[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        xs[j] = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) x[j][g] = 0.0;
    }
    int issued = i0;  // next op whose operands have not been requested yet
    int4 n0 = __ldg(reinterpret_cast<const int4*>(a.ops + i0));
    int4 n1 = __ldg(reinterpret_cast<const int4*>(a.ops + i0) + 1);
    double2 qn1 = __ldg(a.qfac + (size_t)n0.w * SV_KMAX + k);
    double2 qn2 = (n1.y & SV_OP_UNARY) ? make_double2(0.0, 0.0) : __ldg(a.qfac + (size_t)n1.x * SV_KMAX + k);

    for (int i = i0; i < i1; i++) {
        // ops before i are complete (stored, fenced where somebody reads them): request what has become possible
        while (issued < i1 && issued < i + STAGES && ready_after(issued) < i) {
            issue(issued);
            issued++;
        }
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
                if (flags & SV_OP_LEAF1) {
                    sv_lds_slice(base + loff[j], half0, x[j]);
                    xs[j] = 0.0;
                } else {
                    sv_lds_slice(base + voff[j], half0, x[j]);
                    xs[j] = sv_lds64(base + 4096 + soff[j]);
                }
            }
            if (!unary) {
                if (flags & SV_OP_LEAF2) {
                    sv_lds_slice(base + 2048 + loff[j], half0, w[j]);
                } else {
                    sv_lds_slice(base + 2048 + voff[j], half0, w[j]);
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
        if (flags & SV_OP_FENCE) {
            // a later op of this launch reads what was just stored through the copy engine
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp();
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
