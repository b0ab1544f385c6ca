// ssimu2_finalize.cuh — K6: fixed-order reduction of the per-CTA partial sums, the 108-weight
// sum and the final nonlinear map, one CTA per candidate.
//
// Determinism: every (scale, channel, statistic) is summed by one warp, lane l taking CTAs
// l, l+32, ... in order, then a fixed shuffle tree — no atomics, so a score is a pure function
// of the input bytes and the launch geometry.
#pragma once

#include "ssimu2_common.cuh"

namespace oavif {

struct FinalArgs {
    int n_scales;
    int w[kMaxScales], h[kMaxScales];
    int first_cta[kMaxScales + 1];  // per scale: 3 channels x ctas_per_channel
    int ctas_per_channel[kMaxScales];
    const double *partials;         // [candidate][cta][6]
    long long partials_stride;
    double *sums;                   // [candidate][6][18]
    double *scores;                 // [candidate]
    // 0: weight i belongs to slot ((c*6 + scale)*2 + n)*3 + k whether or not the scale exists (missing scales
    //    contribute zero); 1: the published loop's running index over the scales PRESENT
    //    (`scale < scales.size()`, i++): slot ((c*n_scales + scale)*2 + n)*3 + k.  Identical for six scales.
    int contiguous_weights;
};

__constant__ double c_weights[108] = {
    0.0, 0.0007376606707406586, 0.0, 0.0, 0.0007793481682867309, 0.0, 0.0,
    0.0004371155730107379, 0.0, 1.1041726426657346, 0.00066284834129271,
    0.00015231632783718752, 0.0, 0.0016406437456599754, 0.0, 1.8422455520539298,
    11.441172603757666, 0.0, 0.0007989109436015163, 0.000176816438078653, 0.0,
    1.8787594979546387, 10.94906990605142, 0.0, 0.0007289346991508072,
    0.9677937080626833, 0.0, 0.00014003424285435884, 0.9981766977854967,
    0.00031949755934435053, 0.0004550992113792063, 0.0, 0.0, 0.0013648766163243398,
    0.0, 0.0, 0.0, 0.0, 0.0, 7.466890328078848, 0.0, 17.445833984131262,
    0.0006235601634041466, 0.0, 0.0, 6.683678146179332, 0.00037724407979611296,
    1.027889937768264, 225.20515300849274, 0.0, 0.0, 19.213238186143016,
    0.0011401524586618361, 0.001237755635509985, 176.39317598450694, 0.0, 0.0,
    24.43300999870476, 0.28520802612117757, 0.0004485436923833408, 0.0, 0.0, 0.0,
    34.77906344483772, 44.835625328877896, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
    0.0008680556573291698, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0005313191874358747, 0.0,
    0.00016533814161379112, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0004179171803251336,
    0.0017290828234722833, 0.0, 0.0020827005846636437, 0.0, 0.0, 8.826982764996862,
    23.19243343998926, 0.0, 95.1080498811086, 0.9863978034400682, 0.9834382792465353,
    0.0012286405048278493, 171.2667255897307, 0.9807858872435379, 0.0, 0.0, 0.0,
    0.0005130064588990679, 0.0, 0.00010854057858411537};

// grid = n_candidates, block = 1024 (32 warps).
__global__ void __launch_bounds__(1024) k_finalize(const __grid_constant__ FinalArgs a)
{
    __shared__ double s_sum[kMaxScales * 18];
    const int cand = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *P = a.partials + (long long)cand * a.partials_stride;

    for (int i = threadIdx.x; i < kMaxScales * 18; i += blockDim.x) s_sum[i] = 0.0;
    __syncthreads();

    const int nslots = a.n_scales * 18;  // slot = scale*18 + channel*6 + statistic
    for (int slot = warp; slot < nslots; slot += 32) {
        const int s = slot / 18, r = slot - s * 18;
        const int c = r / 6, j = r - c * 6;
        const int n = a.ctas_per_channel[s];
        const double *q = P + ((long long)a.first_cta[s] + (long long)c * n) * 6 + j;
        double acc = 0.0;
        for (int i = lane; i < n; i += 32) acc += q[(long long)i * 6];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) s_sum[slot] = acc;
    }
    __syncthreads();

    double *out = a.sums + (long long)cand * kMaxScales * 18;
    for (int i = threadIdx.x; i < kMaxScales * 18; i += blockDim.x) out[i] = s_sum[i];

    // the 108 weighted terms in parallel (each is a couple of binary64 square roots), then ONE thread adds
    // them in index order, so the sum is the same sequence of additions as a serial evaluation
    __shared__ double s_term[108];
    if (threadIdx.x < 108) {
        const int i = threadIdx.x;           // i = ((c*6 + s)*2 + n)*3 + k
        const int k = i % 3, n = (i / 3) & 1, s = (i / 6) % kMaxScales, c = i / (6 * kMaxScales);
        double term = 0.0;
        if (s < a.n_scales) {
            const double opp = 1.0 / ((double)a.w[s] * (double)a.h[s]);
            const double *v = s_sum + s * 18 + c * 6;
            // n == 0: 1-norm averages; n == 1: 4-norm = (mean of 4th powers)^(1/4)
            const double e = n ? sqrt(sqrt(opp * v[2 * k + 1])) : opp * v[2 * k];
            const int wi = a.contiguous_weights ? ((c * a.n_scales + s) * 2 + n) * 3 + k : i;
            term = c_weights[wi] * fabs(e);
        }
        s_term[i] = term;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ssim = 0.0;
        for (int i = 0; i < 108; ++i) ssim += s_term[i];
        ssim = ssim * 0.9562382616834844;
        ssim = 2.326765642916932 * ssim - 0.020884521182843837 * ssim * ssim +
               6.248496625763138e-05 * ssim * ssim * ssim;
        a.scores[cand] = (ssim > 0.0) ? 100.0 - 10.0 * pow(ssim, 0.6276336467831387) : 100.0;
    }
}

}  // namespace oavif
