// walk.cuh -- the default pruning kernel: single-launch walk over any number of dependent ops, Q-specialised contraction,
// variant partials only (constant-site partials are handled algebraically, core_planner.cuh), root/constant genotype 0.
//
// Same arithmetic, same layouts, same op descriptors and the same bits as k_prune<false, true, true> (prune.cuh), which
// stays for the level-batched schedule, the generic matrices and the per-site constant-site twin.  This kernel exists
// because the walk is bound by instruction issue and exposed latency, not by DRAM (profiles/README.md): it does the same
// work in ~40 % fewer instructions per op --
//   * one address computation and ONE 256-bit load per child whatever it is: base pointer, shift and category offset are
//     selected from the op's (uniform) flags, so the leaf / stored-partial distinction costs no branch;
//   * the product of the two children's contributions is formed in place in the registers that carry the result to the
//     next op (no copies), the second child's block is skipped by one uniform branch for the unary root;
//   * binary exponents are integers and the power-of-two renormalisation only runs where the host's magnitude bounds ask
//     for it (SV_OP_RENORM);
//   * leaf operands of the op a few steps ahead are asked into L2 with one predicated prefetch per operand.
#pragma once
#include "prune.cuh"

// base + row * stride with ONE wide multiply-add (IMAD.WIDE.U32); written as a shift the compiler spends four
// instructions per address (two funnel shifts, add, add-with-carry)
__device__ __forceinline__ const char* sv_row_addr(const char* base, unsigned row, unsigned stride) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(row), "r"(stride), "l"((unsigned long long)base));
    return reinterpret_cast<const char*>(r);
}

// binary exponents through addresses formed by sv_row_addr: ptxas cannot see that they are global and emits generic LD / ST;
// -DSV_WALK_GLOBAL_SCALE (experiment) states the address space
__device__ __forceinline__ int sv_ld_scale(const char* p) {
#ifdef SV_WALK_GLOBAL_SCALE
    int v;
    asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return *reinterpret_cast<const int*>(p);
#endif
}
__device__ __forceinline__ void sv_st_scale(const char* p, int v) {
#ifdef SV_WALK_GLOBAL_SCALE
    asm volatile("st.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    *reinterpret_cast<int*>(const_cast<char*>(p)) = v;
#endif
}

// the 32-byte op descriptor: one 256-bit load (-DSV_WALK_DESC256, experiment) or two 128-bit ones
__device__ __forceinline__ void sv_ld_desc(const SvOp* p, int4& lo, int4& hi) {
#ifdef SV_WALK_DESC256
    unsigned long long r0, r1, r2, r3;
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r0), "=l"(r1), "=l"(r2), "=l"(r3) : "l"(p));
    lo = make_int4((int)(unsigned)r0, (int)(unsigned)(r0 >> 32), (int)(unsigned)r1, (int)(unsigned)(r1 >> 32));
    hi = make_int4((int)(unsigned)r2, (int)(unsigned)(r2 >> 32), (int)(unsigned)r3, (int)(unsigned)(r3 >> 32));
#else
    const int4* q = reinterpret_cast<const int4*>(p);
    lo = __ldg(q);
    hi = __ldg(q + 1);
#endif
}

template <int MINB>
__global__ void __launch_bounds__(SV_PRUNE_THREADS, sv_prune_min_blocks(MINB ? MINB : SV_PRUNE_MINB_Q)) k_walk(SvPruneArgs a) {
    constexpr int NS = SV_PRUNE_NS_Q;
    constexpr int SITES_PER_BLOCK = (SV_PRUNE_THREADS / 4) * NS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int klane = lane & 3;
    const int k = min(klane, a.K - 1);
    const int site0 = blockIdx.x * SITES_PER_BLOCK + warp * (8 * NS) + (lane >> 2);
    // lanes beyond the last site (category) are clamped onto the last valid one: they compute duplicates, only their
    // stores are suppressed (see k_prune)
    bool site_ok[NS];
    unsigned srow[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const int site = site0 + 8 * j;
        site_ok[j] = site < a.S;
        srow[j] = (unsigned)(site_ok[j] ? site : a.S - 1);
        asm volatile("" : "+r"(srow[j]));  // keep it: re-deriving it costs three instructions per use
    }
    const char* const leaf_base = reinterpret_cast<const char*>(a.leafX);
    const char* const part_base = reinterpret_cast<const char*>(a.part) + k * 32;
    const char* const scale_base = reinterpret_cast<const char*>(a.scale);
    // prefetch role of this lane: the lanes with category index j < NS ask for the line of their site of pass j
    unsigned pf_srow = srow[0];
#pragma unroll
    for (int j = 1; j < NS; j++) pf_srow = klane == j ? srow[j] : pf_srow;
    const bool pf_lane = klane < NS;

    const int i0 = a.op_begin, i1 = a.op_end;
    if (a.pf_dist > 0 && pf_lane) {
        // the first pf_dist ops of the list: nobody could announce their leaf operands earlier
        const int jn = min(i1, i0 + a.pf_dist);
#pragma unroll 1
        for (int j = i0; j < jn; j++) {
            const int4 d0 = __ldg(reinterpret_cast<const int4*>(a.ops + j));
            const int fl = __ldg(reinterpret_cast<const int*>(a.ops + j) + 5);
            if (fl & SV_OP_FAR1) sv_prefetch_leaf(leaf_base + ((size_t)((unsigned)d0.y + pf_srow) << 5));
            if (fl & SV_OP_FAR2) sv_prefetch_leaf(leaf_base + ((size_t)((unsigned)d0.z + pf_srow) << 5));
        }
    }

    double x[NS][4];  // the previous op's result, consumed from registers by an op that has it as child 1 (SV_OP_FWD1)
    int xe[NS];       // its binary exponent
#pragma unroll
    for (int j = 0; j < NS; j++) {
        xe[j] = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) x[j][g] = 0.0;
    }
    int4 n0 = make_int4(0, 0, 0, 0), n1 = make_int4(0, 0, 0, 0);
    if (i0 < i1) sv_ld_desc(a.ops + i0, n0, n1);
#pragma unroll 1
    for (int i = i0; i < i1; i++) {
        __syncwarp();  // scratch slots are reused (streamed mode): keep the warp's reads of an op ahead of the next op's writes
        const int4 o0 = n0, o1 = n1;
        if (i + 1 < i1) sv_ld_desc(a.ops + i + 1, n0, n1);  // the next op's descriptor, one op ahead: off the chain descriptor -> addresses -> loads
        const int flags = o1.y;
        // leaf operands of the op pf_dist ahead: into L2 now
        if ((flags & SV_OP_PF1) && pf_lane) sv_prefetch_leaf(leaf_base + ((size_t)((unsigned)o1.z + pf_srow) << 5));
        if ((flags & SV_OP_PF2) && pf_lane) sv_prefetch_leaf(leaf_base + ((size_t)((unsigned)o1.w + pf_srow) << 5));

        const bool has2 = !(flags & SV_OP_UNARY);
        const bool cz0 = flags & SV_OP_ZEROD;
        SV_BOUND(o0.w >= 0 && o0.w < a.n_matrix_slots);
        const double2 q1 = __ldg(a.qfac + (size_t)o0.w * SV_KMAX + k);
        double2 q2;
        double w[NS][4];
        int we[NS];
        // ---- every global load of the op up front ----
        if (has2) {
            q2 = __ldg(a.qfac + (size_t)o1.x * SV_KMAX + k);
            const bool leaf2 = flags & SV_OP_LEAF2;
            const char* const base = leaf2 ? leaf_base : part_base;
            const unsigned stride = leaf2 ? 32u : 128u;  // one wide multiply-add forms the address
#pragma unroll
            for (int j = 0; j < NS; j++) {
                const unsigned row = (unsigned)o0.z + srow[j];
                SV_BOUND(row < (leaf2 ? a.leaf_rows : a.part_rows) && o1.x >= 0 && o1.x < a.n_matrix_slots);
                sv_ld256(reinterpret_cast<const double*>(sv_row_addr(base, row, stride)), w[j]);
                we[j] = 0;
                if (!leaf2) we[j] = sv_ld_scale(sv_row_addr(scale_base, row, 4u));
            }
        }
        if (!(flags & SV_OP_FWD1)) {
            const bool leaf1 = flags & SV_OP_LEAF1;
            const char* const base = leaf1 ? leaf_base : part_base;
            const unsigned stride = leaf1 ? 32u : 128u;
#pragma unroll
            for (int j = 0; j < NS; j++) {
                const unsigned row = (unsigned)o0.y + srow[j];
                SV_BOUND(row < (leaf1 ? a.leaf_rows : a.part_rows));
                sv_ld256(reinterpret_cast<const double*>(sv_row_addr(base, row, stride)), x[j]);
                xe[j] = 0;
                if (!leaf1) xe[j] = sv_ld_scale(sv_row_addr(scale_base, row, 4u));
            }
        }
        // ---- contraction(s) and product; both arms of the (rare) renormalisation write the registers that carry the
        //      result to the next op, so no copies are needed where they meet ----
        double y[NS][4];
        sv_matvec_q<NS>(q1, x, y, cz0);
        if (has2) {
            double z[NS][4];
            sv_matvec_q<NS>(q2, w, z, cz0);
#pragma unroll
            for (int j = 0; j < NS; j++) {
#pragma unroll
                for (int p = 0; p < 4; p++) y[j][p] *= z[j][p];
                xe[j] += we[j];
            }
        }
        if (flags & SV_OP_RENORM) {
#pragma unroll
            for (int j = 0; j < NS; j++) {
                int hi = sv_maxhi4(y[j]);
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
                int sh;
                const double f = sv_pow2_rescale_hi_int(hi, &sh);
#pragma unroll
                for (int p = 0; p < 4; p++) x[j][p] = y[j][p] * f;
                xe[j] += sh;
            }
        } else {
#pragma unroll
            for (int j = 0; j < NS; j++)
#pragma unroll
                for (int p = 0; p < 4; p++) x[j][p] = y[j][p];
        }
        if (!(flags & SV_OP_NOSTORE)) {
#pragma unroll
            for (int j = 0; j < NS; j++) {
                if (!site_ok[j]) continue;  // clamped duplicate lanes never store
                const unsigned row = (unsigned)o0.x + srow[j];
                SV_BOUND(row < a.part_rows);
                sv_st256(reinterpret_cast<double*>(const_cast<char*>(sv_row_addr(part_base, row, 128u))), x[j]);
                sv_st_scale(sv_row_addr(scale_base, row, 4u), xe[j]);  // every category lane writes the (identical) exponent
            }
        }
        if (flags & SV_OP_ROOT) {
            // integrate over categories and take the log (ScsBeerLikelihoodCore.java:153-218, 255-269); the scales of ALL
            // leaves, which the walk never carried, come in here
#pragma unroll
            for (int j = 0; j < NS; j++) {
                double v = (klane < a.K ? a.prop[klane] : 0.0) * x[j][0];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (site_ok[j] && klane == 0) {
                    const long long s = a.out_offset + site0 + 8 * j;
                    a.site_logl[s] = log(v) + fma((double)xe[j], SV_LN2, __ldg(a.lsum + s));
                }
            }
        }
    }
}
