// prune_rp.cuh -- the Q-specialised, variant-only pruning walk as a two-stage software pipeline in registers.
//
// Same arithmetic, op list and results (bit for bit) as k_prune<false, true, true>.  That kernel is latency bound: per
// op a warp issues its operand loads, waits for DRAM, computes ~250 instructions and only then starts the next op.  In
// a post-order walk the first operand of most ops is the previous result (SV_OP_FWD1, in registers), so what an op
// waits for is its SECOND child.  Here the second child of op i + 1 is loaded into a second register set BEFORE op i is
// computed -- and before op i waits for its own second child, which was requested one op earlier: two loads are in
// flight per lane alternately and an op costs about (latency + compute) / 2 instead of their sum.  The two sets swap
// roles by unrolling the walk by two (no register moves).  An op whose second child is the slot the previous op is
// about to write (never the case for the planner's post-order lists, possible with reused scratch slots) loads it late,
// like the plain walk; a first child that does not come from registers (both children leaves) is loaded at the top of
// its op as before.  20 more registers than the plain walk: 5 blocks of 128 threads per SM instead of 6.
//
// MEASURED (B200, 1 000 cells x 100 000 sites): 4.17 ms at 5 blocks/SM (96 registers, 52 bytes of spill) and 4.01 ms at
// 4 blocks/SM (120 registers, no spill) against 2.50 ms for the plain walk -- SLOWER, bit-identical
// (tests/test_gpu_parity.py::test_pipelined_walk_is_bit_identical), so the kernel stays opt-in (SIEVE_B200_RP=1 / 4) as
// the record of the experiment.  Together with two other measurements -- storing a third as many partials
// (SIEVE_B200_TRANSIENT_LEAVES=32: 5.5 -> 2 GB of writes) changes the walk by 1 %, and fetching the 32-byte descriptor
// one op ahead gains 7 % -- this says that the walk is bound neither by DRAM bandwidth nor simply by the load latency
// of one warp -- OR that the pipeline never formed: decoding the control fields of the SASS (bits 105..125 of each
// instruction) shows that ptxas puts EVERY global load of this kernel on the same scoreboard (write barrier 5), and the
// first arithmetic instruction of an op waits on that scoreboard (wait mask 100111): it waits for the prefetch it has
// just issued as well, so the two loads are serialised after all and only the occupancy was lost.  CUDA C++ offers no
// handle on scoreboard assignment; a pipeline that ptxas cannot merge needs a different completion mechanism
// (cp.async groups -- prune_pipe.cuh, which pays with a second pass through L1TEX -- or TMA with an mbarrier).
// A first version (earlier the same day) moved the prefetched operands into the working registers BEFORE issuing the
// next prefetch -- the move waits for the data, so only one op's loads were ever in flight -- 3.37 vs 2.85 ms.
#pragma once
#include "prune.cuh"

template <bool GENO0, int MINB>
__global__ void __launch_bounds__(SV_PRUNE_THREADS, MINB) k_prune_rp(SvPruneArgs a) {
    constexpr int NS = 2;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int k = min(lane & 3, a.K - 1);
    const int klane = lane & 3;
    const int site0 = blockIdx.x * ((SV_PRUNE_THREADS / 4) * NS) + warp * (8 * NS) + (lane >> 2);
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

    double x[NS][4], xs[NS];  // the running result (forwarded to the next op when it asks for it)
    double wA[NS][4], wsA[NS], wB[NS][4], wsB[NS];  // the two register sets for second children
#pragma unroll
    for (int j = 0; j < NS; j++) {
        xs[j] = wsA[j] = wsB[j] = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) x[j][g] = wA[j][g] = wB[j][g] = 0.0;
    }
    // requests the second child of the op described by (d0, flags); nothing here waits
    auto issue_child2 = [&](const int4& d0, const int flags, double (&w)[NS][4], double (&ws)[NS]) {
#pragma unroll
        for (int j = 0; j < NS; j++) {
            const size_t row = (unsigned)d0.z + srow[j];
            if (flags & SV_OP_LEAF2) {
                sv_ld256_nc(a.leafX + row * 4, w[j]);
                ws[j] = 0.0;
            } else {
                sv_ld256(a.part + row * 16 + k * 4, w[j]);
                ws[j] = a.scale[row];
            }
        }
    };
    int4 n0 = __ldg(reinterpret_cast<const int4*>(a.ops + i0));
    int4 n1 = __ldg(reinterpret_cast<const int4*>(a.ops + i0) + 1);
    bool haveA = false, haveB = false;
    if (!(n1.y & SV_OP_UNARY)) {
        issue_child2(n0, n1.y, wA, wsA);
        haveA = true;
    }

    // one op: operands of its second child in (w, ws) if `have`; the next op's second child goes to (wn, wsn)
    auto step = [&](const int i, double (&w)[NS][4], double (&ws)[NS], const bool have, double (&wn)[NS][4],
                    double (&wsn)[NS], bool& have_n) {
        __syncwarp();  // scratch slots are reused (streamed / sparse walks): keep the warp's reads ahead of the next op's writes
        const int4 o0 = n0, o1 = n1;
        if (i + 1 < i1) {
            n0 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1));
            n1 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1) + 1);
        }
        const unsigned dst_base = (unsigned)o0.x, c1_base = (unsigned)o0.y;
        const int m1 = o0.w, m2 = o1.x, flags = o1.y;
        const bool unary = flags & SV_OP_UNARY;
        const double2 q1 = __ldg(a.qfac + (size_t)m1 * SV_KMAX + k);
        double2 q2 = make_double2(0.0, 0.0);
        if (!unary) q2 = __ldg(a.qfac + (size_t)m2 * SV_KMAX + k);
        if (!(flags & SV_OP_FWD1)) {
#pragma unroll
            for (int j = 0; j < NS; j++) {
                const size_t c1_row = c1_base + srow[j];
                if (flags & SV_OP_LEAF1) {
                    sv_ld256_nc(a.leafX + c1_row * 4, x[j]);
                    xs[j] = 0.0;
                } else {
                    sv_ld256(a.part + c1_row * 16 + k * 4, x[j]);
                    xs[j] = a.scale[c1_row];
                }
            }
        }
        if (!unary && !have) issue_child2(o0, flags, w, ws);  // could not be requested early
        // the next op's second child: requested now, in flight while this op waits for its own and computes
        have_n = false;
        if (i + 1 < i1) {
            const int nf = n1.y;
            const bool dep = !(nf & SV_OP_LEAF2) && (unsigned)n0.z == dst_base;  // it is what this op is about to write
            if (!(nf & SV_OP_UNARY) && !dep) {
                issue_child2(n0, nf, wn, wsn);
                have_n = true;
            }
        }
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
            xs[j] = (xs[j] + (unary ? 0.0 : ws[j])) + sh;
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
    };
    for (int i = i0; i < i1; i += 2) {
        step(i, wA, wsA, haveA, wB, wsB, haveB);
        if (i + 1 < i1) step(i + 1, wB, wsB, haveB, wA, wsA, haveA);
    }
}
