// common.cuh -- small device helpers shared by the kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Bounds-checked build (make variant NAME=checked DEFS=-DSV_CHECK_BOUNDS; never the product library): every pool and table
// index of the hot kernels is asserted on the device.  compute-sanitizer is closed on the GPU pool this was developed on.
#ifdef SV_CHECK_BOUNDS
#include <cstdio>
#define SV_BOUND(cond)                                                                      \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            printf("SV_BOUND failed: %s (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
            __trap();                                                                       \
        }                                                                                   \
    } while (0)
#else
#define SV_BOUND(cond) ((void)0)
#endif

#define SV_MIN_VALUE 4.9406564584124654e-324  // java.lang.Double.MIN_VALUE
#define SV_LN2 0.6931471805599453094
#define SV_G 4                                 // genotype states (ScsFiniteMuExtendedModel.java:31-34)
#define SV_KMAX 4                              // site-rate categories handled per site by 4 adjacent lanes
#define SV_MAT_STRIDE 18                       // doubles between the K matrices in shared memory (bank padding)

// One post-order step of the pruning: parent <- (child1, child2).  Rows index the device pools: row = slot * S, so that
// the address of (slot, site) is base + (row + site) * stride with no multiplication in the kernel.  The host puts a
// child that is the destination of the immediately preceding op FIRST and flags it SV_OP_FWD1 (the product of the two
// children's contributions commutes), so the walk can consume it from registers.
struct __align__(16) SvOp {
    unsigned dst_row;  // internal slot (buffer * n_internal + internal index) * S
    unsigned c1_row;   // child 1: leaf slot (buffer * n_leaves + leaf) * S  or  internal slot * S
    unsigned c2_row;   // child 2, unused for a unary op
    int m1, m2;        // matrix slots (buffer * n_nodes + node) of the two children
    int flags;         // SV_OP_*
    unsigned pf_row1, pf_row2;  // see SV_OP_PF*
};  // 32 bytes: two 128-bit loads.  Rows are 32-bit: every pool holds fewer than 2^32 (slot, site) rows (checked at create).
#define SV_OP_LEAF1 1
#define SV_OP_LEAF2 2
#define SV_OP_UNARY 4
#define SV_OP_ROOT 8
#define SV_OP_FWD1 16
#define SV_OP_ZEROD 64    // one of the two distances is exactly 0: take the Double.MIN_VALUE clamp path
// L2 prefetch annotations (sv_link_prefetch): op i carries, in pf_row1 / pf_row2, the rows of the LEAF operands that op
// i + D will read (leaf partials always come from DRAM: every one is read once per evaluation); the walk asks for
// their lines a few ops ahead so that the op that consumes them pays an L2 hit instead of a DRAM round trip.
#define SV_OP_PF1 256
#define SV_OP_PF2 1024
#define SV_OP_FAR1 4096   // this op's own child 1 / child 2 is such an operand (used for the first D ops of a list,
#define SV_OP_FAR2 8192   // which no earlier op could announce)
#define SV_OP_RENORM 16384  // renormalise this op's result by a power of two (sv_bound_op decides; the root always does)
#define SV_OP_NOSTORE 32  // the result is consumed from registers by the next op and never needed again: skip the writes

// 256-bit global accesses (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a): one (site, category) state vector.
__device__ __forceinline__ void sv_ld256(const double* p, double (&v)[4]) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
// read-only variant for data that no kernel in flight writes (leaf partials during pruning, counts)
__device__ __forceinline__ void sv_ld256_nc(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void sv_st256(double* p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

// one line into L2: no destination register, no scoreboard entry
__device__ __forceinline__ void sv_prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// the walk's announcement of a leaf operand: L2 by default, L1 with -DSV_PF_L1 (experiment)
__device__ __forceinline__ void sv_prefetch_leaf(const void* p) {
#ifdef SV_PF_L1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

__device__ __forceinline__ double sv_max4(const double (&v)[4]) { return fmax(fmax(v[0], v[1]), fmax(v[2], v[3])); }

// For NON-NEGATIVE doubles the order of the values is the order of their high words as signed integers, and the
// renormalisation only needs the exponent of the maximum: 3 integer max instead of 3 double compares + selects.
// (NaN has the largest high word, so it surfaces as "exponent 0x7ff" and the values are left alone.)
__device__ __forceinline__ int sv_maxhi4(const double (&v)[4]) {
    return max(max(__double2hiint(v[0]), __double2hiint(v[1])), max(__double2hiint(v[2]), __double2hiint(v[3])));
}
// factor = 2^-e and *shift_ln = e ln 2 from the high word of the maximum (see sv_pow2_rescale); selects only, no branch
__device__ __forceinline__ double sv_pow2_rescale_hi(int hi, double* shift_ln) {
    const int e = (hi >> 20) & 0x7ff;
    const bool normal = hi > 0 && (unsigned)(e - 1) < 2045u;
    const bool sub = hi > 0 && e == 0;  // subnormal maximum with a non-zero high word (Double.MIN_VALUE clamps only)
    const int fe = normal ? 2046 - e : (sub ? 2045 : 1023);
    const int se = normal ? e - 1023 : (sub ? -1022 : 0);
    *shift_ln = (double)se * SV_LN2;  // zero / tiny subnormal / inf / NaN: left alone
    return __hiloint2double(fe << 20, 0);
}

// the same with the exponent as an integer (the walk keeps a node's scale as an exact sum of binary exponents)
__device__ __forceinline__ double sv_pow2_rescale_hi_int(int hi, int* e_out) {
    const int e = (hi >> 20) & 0x7ff;
    const bool normal = hi > 0 && (unsigned)(e - 1) < 2045u;
    const bool sub = hi > 0 && e == 0;  // subnormal maximum with a non-zero high word (Double.MIN_VALUE clamps only)
    const int fe = normal ? 2046 - e : (sub ? 2045 : 1023);
    *e_out = normal ? e - 1023 : (sub ? -1022 : 0);  // zero / tiny subnormal / inf / NaN: left alone
    return __hiloint2double(fe << 20, 0);
}

// Exact power-of-two renormalisation: returns factor = 2^-e and *shift_ln = e * ln 2 with e = floor(log2 mx),
// so that mx * factor is in [1, 2).  Zero / non-finite maxima are left alone.
__device__ __forceinline__ double sv_pow2_rescale(double mx, double* shift_ln) {
    const int hi = __double2hiint(mx);
    const int e = (hi >> 20) & 0x7ff;
    if (hi > 0 && e >= 1 && e <= 2045) {
        *shift_ln = (double)(e - 1023) * SV_LN2;
        return __hiloint2double((2046 - e) << 20, 0);
    }
    if (mx > 0.0 && e == 0) {  // subnormal maximum (only reachable through the Double.MIN_VALUE clamps)
        *shift_ln = -1022.0 * SV_LN2;
        return __hiloint2double(2045 << 20, 0);  // 2^1022
    }
    *shift_ln = 0.0;
    return 1.0;
}
