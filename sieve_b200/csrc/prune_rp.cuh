// prune_rp.cuh -- the Q-specialised, variant-only pruning walk with the NEXT op's operands prefetched into registers.
//
// Same arithmetic, op list and results (bit for bit) as k_prune<false, true, true>.  That kernel is latency bound
// (profiles/r01g_*: issue slots 50 % busy, long-scoreboard stalls, DRAM 45 %): a warp issues the loads of an op, waits
// ~1 us for DRAM, computes ~250 instructions, and only then learns the addresses of the next op.  Here the loads of op
// i + 1 (both children's vectors and scales for the lane's two sites, the two branch factors) are issued BEFORE op i is
// computed and land in a second set of registers; op i + 1 starts with its operands present.  More registers (4 blocks
// of 128 threads per SM instead of 6), but every resident warp now always has a full op of loads in flight.
// The staging through shared memory that prune_pipe.cuh tried for the same purpose doubled the L1TEX traffic; registers
// do not.
//
// MEASURED (B200, 1 000 cells x 100 000 sites, gpurun 2026-10-20): 3.37 ms against 2.85 ms for the plain walk -- SLOWER,
// so this kernel too is opt-in (SIEVE_B200_RP=1) and kept as the record of the experiment: 124 registers leave 16 warps
// per SM instead of 24, and the extra warps of the plain kernel hide more latency than one op of look-ahead per warp.
// Results are bit-identical (tests/test_gpu_parity.py::test_pipelined_walk_is_bit_identical).
#pragma once
#include "prune.cuh"

#ifndef SV_RP_MINB
#define SV_RP_MINB 4
#endif

struct SvRpOperands {
    double x[2][4], w[2][4];  // child 1 / child 2 vectors of the lane's two sites
    double xs[2], ws[2];      // their scales
    double2 q1, q2;           // branch factors
    unsigned dst_base;
    int flags;
};

// issue the loads of op `i`; nothing here waits
template <int NS>
__device__ __forceinline__ void sv_rp_issue(const SvPruneArgs& a, int i, int k, const unsigned (&srow)[NS], SvRpOperands& o) {
    const int4* opp = reinterpret_cast<const int4*>(a.ops + i);
    const int4 o0 = __ldg(opp), o1 = __ldg(opp + 1);
    const unsigned c1_base = (unsigned)o0.y, c2_base = (unsigned)o0.z;
    const int m1 = o0.w, m2 = o1.x, flags = o1.y;
    o.dst_base = (unsigned)o0.x;
    o.flags = flags;
    o.q1 = __ldg(a.qfac + (size_t)m1 * SV_KMAX + k);
    o.q2 = make_double2(0.0, 0.0);
    if (!(flags & SV_OP_UNARY)) o.q2 = __ldg(a.qfac + (size_t)m2 * SV_KMAX + k);
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const size_t c1_row = c1_base + srow[j], c2_row = c2_base + srow[j];
        if (!(flags & SV_OP_FWD1)) {
            if (flags & SV_OP_LEAF1) {
                sv_ld256_nc(a.leafX + c1_row * 4, o.x[j]);
                o.xs[j] = 0.0;
            } else {
                sv_ld256(a.part + c1_row * 16 + k * 4, o.x[j]);
                o.xs[j] = a.scale[c1_row];
            }
        }
        o.ws[j] = 0.0;
        if (!(flags & SV_OP_UNARY)) {
            if (flags & SV_OP_LEAF2) {
                sv_ld256_nc(a.leafX + c2_row * 4, o.w[j]);
                o.ws[j] = 0.0;
            } else {
                sv_ld256(a.part + c2_row * 16 + k * 4, o.w[j]);
                o.ws[j] = a.scale[c2_row];
            }
        }
    }
}

template <bool GENO0>
__global__ void __launch_bounds__(SV_PRUNE_THREADS, SV_RP_MINB) k_prune_rp(SvPruneArgs a) {
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
#pragma unroll
    for (int j = 0; j < NS; j++) {
        xs[j] = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) x[j][g] = 0.0;
    }
    SvRpOperands nx;
#pragma unroll
    for (int j = 0; j < NS; j++)
#pragma unroll
        for (int g = 0; g < 4; g++) nx.x[j][g] = nx.w[j][g] = 0.0;
    nx.xs[0] = nx.xs[1] = nx.ws[0] = nx.ws[1] = 0.0;
    sv_rp_issue<NS>(a, i0, k, srow, nx);
    bool refetch = false;
    for (int i = i0; i < i1; i++) {
        __syncwarp();  // scratch slots are reused (streamed / sparse walks): keep the warp's reads ahead of the next op's writes
        if (refetch) sv_rp_issue<NS>(a, i, k, srow, nx);  // its operand was written by the previous op: could not be fetched early
        // take over the prefetched operands
        const int flags = nx.flags;
        const unsigned dst_base = nx.dst_base;
        const bool unary = flags & SV_OP_UNARY;
        const double2 q1 = nx.q1, q2 = nx.q2;
        double w[NS][4], ws[NS];
#pragma unroll
        for (int j = 0; j < NS; j++) {
            if (!(flags & SV_OP_FWD1)) {
#pragma unroll
                for (int g = 0; g < 4; g++) x[j][g] = nx.x[j][g];
                xs[j] = nx.xs[j];
            }
#pragma unroll
            for (int g = 0; g < 4; g++) w[j][g] = nx.w[j][g];
            ws[j] = nx.ws[j];
        }
        // the next op's loads go out now and are in flight during this op's arithmetic -- unless one of its operands is
        // what this op is about to store (the forwarding covers the common case; anything else is fetched afterwards)
        refetch = false;
        if (i + 1 < i1) {
            const int4 n0 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1));
            const int4 n1 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1) + 1);
            const bool dep1 = !(n1.y & (SV_OP_LEAF1 | SV_OP_FWD1)) && (unsigned)n0.y == dst_base;
            const bool dep2 = !(n1.y & (SV_OP_LEAF2 | SV_OP_UNARY)) && (unsigned)n0.z == dst_base;
            refetch = dep1 || dep2;
            if (!refetch) sv_rp_issue<NS>(a, i + 1, k, srow, nx);
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
