// prune_pipe.cuh -- the Q-specialised, variant-only pruning walk with a two-stage software pipeline.
//
// Same arithmetic, same op list and same results (bit for bit) as k_prune<false, true, true> in prune.cuh; what changes
// is WHEN the operands arrive.  The profile of that kernel at 1 000 cells x 100 000 sites (profiles/r01g_*) shows a
// latency-bound walk: issue slots 50 % busy, 5.7 warps stalled on the long scoreboard per issued instruction, DRAM at
// 45 %.  Every op of the walk starts with ~1 us of exposed DRAM latency for its two children, because their addresses
// are known only when the op's descriptor has been read and the registers that could hold a prefetched op are taken.
//
// Here the operands of op i + 1 (two child vectors and their scales per site, the two branch factors) are copied
// global -> shared with cp.async (LDGSTS, no registers, no thread waits) while op i is computed from the buffer that was
// filled one iteration earlier.  Every lane copies and later reads ONLY ITS OWN slots of the buffer, so
// cp.async.wait_group is the only synchronisation: no barrier, no mbarrier.  The walk's forwarding (SV_OP_FWD1: child 1
// is the previous op's result, still in registers) and transient results (SV_OP_NOSTORE) are unchanged.
//
// Shared memory per block (128 lanes, 2 sites per lane): 2 stages x (2 children x 2 sites x 40 B + 2 x 16 B) x 128 =
// 48 KB -> 4 blocks per SM.  Vectors are stored as two 16-byte planes so that the 128-bit shared loads of a warp touch
// consecutive banks.
//
// MEASURED (B200, 1 000 cells x 100 000 sites, gpurun 2026-10-20): 4.25 ms against 2.81 ms for the plain walk -- SLOWER,
// so this kernel is opt-in (SIEVE_B200_PIPE=1) and kept as the record of the experiment.  Why: the plain walk already
// keeps the L1TEX data pipe ~60 % busy with its 256-bit global loads and stores (profiles/r01f_*); staging the operands
// in shared memory sends every operand byte through that pipe twice (LDGSTS in, LDS out), the 16-byte cp.async
// granularity halves the sector efficiency of the L2 requests, and 48 KB of buffers cost a third of the resident warps.
// The latency it hides is worth less than the L1TEX time it adds.  Results are bit-identical
// (tests/test_gpu_parity.py::test_pipelined_walk_is_bit_identical).
#pragma once
#include "prune.cuh"

#define SV_PIPE_NS 2
#define SV_PIPE_SITES ((SV_PRUNE_THREADS / 4) * SV_PIPE_NS)
// per stage: vec planes [child][site][half][lane] 16 B, scales [child][site][lane] 8 B, q [child][lane] 16 B
#define SV_PIPE_VEC_BYTES (2 * SV_PIPE_NS * 2 * SV_PRUNE_THREADS * 16)
#define SV_PIPE_SC_BYTES (2 * SV_PIPE_NS * SV_PRUNE_THREADS * 8)
#define SV_PIPE_Q_BYTES (2 * SV_PRUNE_THREADS * 16)
#define SV_PIPE_STAGE_BYTES (SV_PIPE_VEC_BYTES + SV_PIPE_SC_BYTES + SV_PIPE_Q_BYTES)
#define SV_PIPE_SMEM (2 * SV_PIPE_STAGE_BYTES)
#ifndef SV_PIPE_MINB
#define SV_PIPE_MINB 4
#endif

__device__ __forceinline__ void sv_cp16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sv_cp8(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sv_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sv_cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void sv_lds_vec(unsigned addr_half0, double (&v)[4]) {
    // half 1 lives SV_PRUNE_THREADS * 16 bytes after half 0
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr_half0));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(addr_half0 + SV_PRUNE_THREADS * 16));
}
__device__ __forceinline__ double sv_lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 sv_lds_d2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

template <bool GENO0>
__global__ void __launch_bounds__(SV_PRUNE_THREADS, SV_PIPE_MINB) k_prune_pipe(SvPruneArgs a) {
    constexpr int NS = SV_PIPE_NS;
    extern __shared__ __align__(16) unsigned char sv_pipe_smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int k = min(lane & 3, a.K - 1);
    const int klane = lane & 3;
    const int site0 = blockIdx.x * SV_PIPE_SITES + warp * (8 * NS) + (lane >> 2);
    bool site_ok[NS];
    unsigned srow[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const int site = site0 + 8 * j;
        site_ok[j] = site < a.S;
        srow[j] = (unsigned)(site_ok[j] ? site : a.S - 1);
    }
    const int rg = GENO0 ? 0 : a.root_genotype;
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(sv_pipe_smem);
    // this lane's slots inside a stage
    const unsigned vec_off = (unsigned)tid * 16;                                   // + ((c * NS + j) * 2 + half) * 2048
    const unsigned sc_off = SV_PIPE_VEC_BYTES + (unsigned)tid * 8;                 // + (c * NS + j) * 1024
    const unsigned q_off = SV_PIPE_VEC_BYTES + SV_PIPE_SC_BYTES + (unsigned)tid * 16;  // + c * 2048

    // issue the copies of op `i` into stage `st`
    auto prefetch = [&](int i, int st) {
        const int4* opp = reinterpret_cast<const int4*>(a.ops + i);
        const int4 o0 = __ldg(opp), o1 = __ldg(opp + 1);
        const unsigned c1_base = (unsigned)o0.y, c2_base = (unsigned)o0.z;
        const int m1 = o0.w, m2 = o1.x, flags = o1.y;
        const unsigned base = smem0 + (unsigned)st * SV_PIPE_STAGE_BYTES;
        sv_cp16(base + q_off, a.qfac + (size_t)m1 * SV_KMAX + k);
        if (!(flags & SV_OP_UNARY)) sv_cp16(base + q_off + SV_PRUNE_THREADS * 16, a.qfac + (size_t)m2 * SV_KMAX + k);
#pragma unroll
        for (int j = 0; j < NS; j++) {
            if (!(flags & SV_OP_FWD1)) {
                const size_t row = c1_base + srow[j];
                const double* src = (flags & SV_OP_LEAF1) ? a.leafX + row * 4 : a.part + row * 16 + k * 4;
                const unsigned d = base + vec_off + (unsigned)((0 * NS + j) * 2) * (SV_PRUNE_THREADS * 16);
                sv_cp16(d, src);
                sv_cp16(d + SV_PRUNE_THREADS * 16, src + 2);
                if (!(flags & SV_OP_LEAF1))
                    sv_cp8(base + sc_off + (unsigned)(0 * NS + j) * (SV_PRUNE_THREADS * 8), a.scale + row);
            }
            if (!(flags & SV_OP_UNARY)) {
                const size_t row = c2_base + srow[j];
                const double* src = (flags & SV_OP_LEAF2) ? a.leafX + row * 4 : a.part + row * 16 + k * 4;
                const unsigned d = base + vec_off + (unsigned)((1 * NS + j) * 2) * (SV_PRUNE_THREADS * 16);
                sv_cp16(d, src);
                sv_cp16(d + SV_PRUNE_THREADS * 16, src + 2);
                if (!(flags & SV_OP_LEAF2))
                    sv_cp8(base + sc_off + (unsigned)(1 * NS + j) * (SV_PRUNE_THREADS * 8), a.scale + row);
            }
        }
        sv_cp_commit();
    };

    const int i0 = a.op_begin, i1 = a.op_end;
    double x[NS][4], xs[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        xs[j] = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) x[j][g] = 0.0;
    }
    if (i0 < i1) prefetch(i0, 0);
    for (int i = i0; i < i1; i++) {
        __syncwarp();  // scratch slots are reused (streamed mode): the warp's reads of an op stay ahead of the next op's writes
        const int st = (i - i0) & 1;
        const int4* opp = reinterpret_cast<const int4*>(a.ops + i);
        const int4 o0 = __ldg(opp), o1 = __ldg(opp + 1);
        const unsigned dst_base = (unsigned)o0.x;
        const int flags = o1.y;
        const bool unary = flags & SV_OP_UNARY;
        // The next op's operands: in flight while this op computes.  An operand that THIS op is about to write cannot
        // be copied early; the host's forwarding covers the only such case (child 1 == this op's result, SV_OP_FWD1),
        // anything else of that kind (never produced by the planner) ends the pipeline for one op.
        bool next_ok = false;
        if (i + 1 < i1) {
            const int4 n0 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1));
            const int4 n1 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1) + 1);
            const bool dep1 = !(n1.y & (SV_OP_LEAF1 | SV_OP_FWD1)) && (unsigned)n0.y == dst_base;
            const bool dep2 = !(n1.y & (SV_OP_LEAF2 | SV_OP_UNARY)) && (unsigned)n0.z == dst_base;
            next_ok = !(dep1 || dep2);
            if (next_ok) prefetch(i + 1, st ^ 1);
        }
        if (next_ok)
            sv_cp_wait<1>();
        else
            sv_cp_wait<0>();
        const unsigned base = smem0 + (unsigned)st * SV_PIPE_STAGE_BYTES;
        const double2 q1 = sv_lds_d2(base + q_off);
        double2 q2 = make_double2(0.0, 0.0);
        double w[NS][4], ws[NS];
#pragma unroll
        for (int j = 0; j < NS; j++) {
            ws[j] = 0.0;
#pragma unroll
            for (int g = 0; g < 4; g++) w[j][g] = 0.0;
        }
        if (!unary) q2 = sv_lds_d2(base + q_off + SV_PRUNE_THREADS * 16);
#pragma unroll
        for (int j = 0; j < NS; j++) {
            if (!(flags & SV_OP_FWD1)) {
                sv_lds_vec(base + vec_off + (unsigned)((0 * NS + j) * 2) * (SV_PRUNE_THREADS * 16), x[j]);
                xs[j] = (flags & SV_OP_LEAF1) ? 0.0 : sv_lds_f64(base + sc_off + (unsigned)(0 * NS + j) * (SV_PRUNE_THREADS * 8));
            }
            if (!unary) {
                sv_lds_vec(base + vec_off + (unsigned)((1 * NS + j) * 2) * (SV_PRUNE_THREADS * 16), w[j]);
                ws[j] = (flags & SV_OP_LEAF2) ? 0.0 : sv_lds_f64(base + sc_off + (unsigned)(1 * NS + j) * (SV_PRUNE_THREADS * 8));
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
        // the pipeline was cut for the next op (its operand is the value just stored): fetch it now
        if (i + 1 < i1 && !next_ok) prefetch(i + 1, st ^ 1);
    }
}
