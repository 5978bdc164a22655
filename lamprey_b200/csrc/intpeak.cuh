// intpeak.cuh -- integer / DPX issue-rate microbenchmark (the roofline of
// this path is the integer ALU / DPX issue rate, which MEASURED_PEAKS.json does not contain).
//
// Every thread runs CH independent dependency chains of one instruction mix; all SMs are
// filled.  The host divides lane-ops by the CUDA-event time.  Modes:
//   0  VIADDMNMX.S16x2            (__viaddmax_s16x2)
//   1  VIMNMX3.S16x2              (__vimax3_s16x2)
//   2  integer add                (two dependent adds per chain step; IADD3 / VIADD, compiler's choice)
//   3  IMAD                       (mad.lo with a run-time multiplier: FMA pipe)
//   4  K1 row mix, compiler add   (VIADDMNMX, VIMNMX3, add, VIADDMNMX: 4 ops per two cells)
//   5  K1 row mix, add as IMAD    (same with the subtraction forced onto the FMA pipe)
//   6  LOP3                       (two dependent 3-input logic ops per chain step)
//   7  VIADDMNMX.S16x2.RELU
//   8  VIMNMX.S16x2 (2-input)
//   9  K1's real inner loop (lsw.cu k_intpeak_k1row: ScoreLane<24>::column2 with its LDS.64 profile loads, K1's launch shape)
#pragma once
#include <stdint.h>

namespace lsw {

constexpr int IP_CH = 8;

template <int MODE>
__global__ void __launch_bounds__(256)
k_intpeak(uint32_t* out, const uint32_t* in, int iters, uint32_t one) {
    uint32_t v[IP_CH], k[IP_CH], m[IP_CH];
    const uint32_t a = in[0], b = in[1], c = in[2];
#pragma unroll
    for (int i = 0; i < IP_CH; i++) { v[i] = in[3] + i + threadIdx.x; k[i] = i; m[i] = in[4] + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < IP_CH; i++) {
            if (MODE == 0) v[i] = __viaddmax_s16x2(v[i], a, b);
            else if (MODE == 1) v[i] = __vimax3_s16x2(v[i], a, b);
            else if (MODE == 2) { v[i] = v[i] + k[i]; k[i] = k[i] + v[i]; }
            else if (MODE == 3) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v[i]) : "r"(v[i]), "r"(one), "r"(a));
            else if (MODE == 4 || MODE == 5) {
                const uint32_t T = __viaddmax_s16x2(m[i], a, v[i]);
                const uint32_t h = __vimax3_s16x2(T, v[i ? i - 1 : IP_CH - 1], c);
                uint32_t u;
                if (MODE == 4) u = h - c;
                else asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(h), "r"(one), "r"(b));
                m[i] = v[i];
                v[i] = u;
                k[i] = __viaddmax_s16x2(h, a, k[i]);
            }
            else if (MODE == 6) { v[i] = (v[i] & k[i]) ^ m[i]; k[i] = (k[i] | v[i]) ^ m[i]; }
            else if (MODE == 7) v[i] = __viaddmax_s16x2_relu(v[i], a, b);
            else if (MODE == 8) v[i] = __vmaxs2(v[i], a) ;
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < IP_CH; i++) s += v[i] + k[i] + m[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

inline int intpeak_ops_per_iter(int mode) {
    return (mode == 4 || mode == 5) ? 4 * IP_CH : (mode == 2 || mode == 6) ? 2 * IP_CH : IP_CH;
}

}  // namespace lsw
