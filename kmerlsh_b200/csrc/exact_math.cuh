// The reference's floating-point formulas, one IEEE binary32 operation per source operation, in
// source order, never fused (the reference's x86-64 build has no FMA; SURVEY.md section 7 hard
// part 2).  Every kernel that decides a merge or writes a centroid uses these and nothing else.
#pragma once
#include <cuda_runtime.h>

// Rows are padded with zeros to a multiple of 4 floats; the padded products are +0 and x + (+0) == x
// (a running sum that starts at +0 is never -0), so walking quads gives the reference's D-term sums.

// sqrt(sum v_i^2), i ascending: the magnitude_* sums of Distance::cosine (function/distance.cc:30-35)
__device__ __forceinline__ float norm_seq(const float4* a, int nq) {
  float s = 0.f;
#pragma unroll 4
  for (int q = 0; q < nq; ++q) {
    const float4 x = a[q];
    s = __fadd_rn(s, __fmul_rn(x.x, x.x));
    s = __fadd_rn(s, __fmul_rn(x.y, x.y));
    s = __fadd_rn(s, __fmul_rn(x.z, x.z));
    s = __fadd_rn(s, __fmul_rn(x.w, x.w));
  }
  return __fsqrt_rn(s);
}

// sum l_i * r_i, i ascending (function/distance.cc:31)
__device__ __forceinline__ float dot_seq(const float4* a, const float4* b, int nq) {
  float s = 0.f;
#pragma unroll 4
  for (int q = 0; q < nq; ++q) {
    const float4 x = a[q], y = b[q];
    s = __fadd_rn(s, __fmul_rn(x.x, y.x));
    s = __fadd_rn(s, __fmul_rn(x.y, y.y));
    s = __fadd_rn(s, __fmul_rn(x.z, y.z));
    s = __fadd_rn(s, __fmul_rn(x.w, y.w));
  }
  return s;
}

// Distance::cosine's return value: 1 - dot / (sqrtf(ml) * sqrtf(mr))  (function/distance.cc:36-37)
__device__ __forceinline__ float cosine_distance(float dot, float nl, float nr) {
  return __fsub_rn(1.f, __fdiv_rn(dot, __fmul_rn(nl, nr)));
}

// p_cluster's test `1 - distance >= threshold` (function/cluster.cc:68-69)
__device__ __forceinline__ bool cos_match(float dot, float nl, float nr, float threshold) {
  return __fsub_rn(1.f, cosine_distance(dot, nl, nr)) >= threshold;
}

// one dimension of AB::SetConsensus (function/funcAB.cc:58-63): cur*c1/all + cand*c2/all, the int
// counts converted like cvtsi2ss (round to nearest even above 2^24)
__device__ __forceinline__ float consensus1(float cur, int c1, float cand, int c2) {
  const float fa = __int2float_rn(c1 + c2);
  const float a = __fdiv_rn(__fmul_rn(cur, __int2float_rn(c1)), fa);
  const float b = __fdiv_rn(__fmul_rn(cand, __int2float_rn(c2)), fa);
  return __fadd_rn(a, b);
}
