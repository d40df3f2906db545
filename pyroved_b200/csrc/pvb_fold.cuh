// Device helpers shared by the latent-side kernels (pvb_latent.cu) and the fused
// small-batch MLP kernels (pvb_mlp.cu): counter-based N(0,1), latent split.
#pragma once
#include "pvb_common.cuh"

namespace pvb {

// ---- Philox4x32-10 ---------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}


// eps for global element index idx at optimizer step `step`: Box-Muller on two 32-bit uniforms
__device__ __forceinline__ float philox_randn(uint64_t idx, uint32_t step, uint64_t seed) {
  uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), step, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float u1 = ((float)r.x + 1.0f) * 2.3283064365386963e-10f;
  float u2 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
  u1 = fminf(fmaxf(u1, 1e-12f), 1.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ---- fold: (phi, dx, dy, s, zc, cond) -> Uv[i] = (U0 | U1 | v) ---------------
struct Split {
  int off_phi, off_t, off_s, off_c;  // offsets into z (-1 if absent)
};
__host__ __device__ inline Split split_of(const pvb_fold_cfg& c) {
  Split s;
  int o = 0;
  s.off_phi = s.off_t = s.off_s = -1;
  if (c.ndim == 1) {
    if (c.inv & PVB_INV_T) { s.off_t = o; o += 1; }
  } else {
    if (c.inv & PVB_INV_R) { s.off_phi = o; o += 1; }
    if (c.inv & PVB_INV_T) { s.off_t = o; o += 2; }
    if (c.inv & PVB_INV_S) { s.off_s = o; o += 1; }
  }
  s.off_c = o;
  return s;
}


// transform parameters of one instance: (cos, sin, scale, dx, dy)  [models/base.py:97-119]
struct Xform { float c, sn, s, dx, dy; };
__device__ __forceinline__ Xform xform_of(const pvb_fold_cfg& cfg, const Split& sp, const float* zi) {
  Xform t;
  t.c = 1.f; t.sn = 0.f; t.s = 1.f; t.dx = 0.f; t.dy = 0.f;
  if (cfg.ndim == 2) {
    if (sp.off_phi >= 0) sincosf(zi[sp.off_phi], &t.sn, &t.c);
    if (sp.off_t >= 0) { t.dx = zi[sp.off_t] * cfg.dx_prior; t.dy = zi[sp.off_t + 1] * cfg.dy_prior; }
    if (sp.off_s >= 0) t.s = 1.f + cfg.sc_prior * zi[sp.off_s];
  } else if (sp.off_t >= 0) {
    t.dx = zi[sp.off_t] * cfg.dx_prior;
  }
  return t;
}
// Uv[.][h] of one instance for hidden unit h  [utils/coord.py:71-75,84-88,60; nets/fc.py:226-235]
__device__ __forceinline__ void fold_unit(const pvb_fold_cfg& cfg, const Split& sp, const Xform& t,
                                          const float* zi, const float* cond_i, const float* Wc,
                                          const float* bc, const float* Wz, int h, float* out) {
  const int LC = cfg.latent_dim + cfg.cond_dim, Hd = cfg.hidden;
  float v = bc[h];
  for (int j = 0; j < cfg.latent_dim; ++j) v = fmaf(Wz[h * LC + j], zi[sp.off_c + j], v);
  for (int j = 0; j < cfg.cond_dim; ++j) v = fmaf(Wz[h * LC + cfg.latent_dim + j], cond_i[j], v);
  if (cfg.ndim == 2) {
    float w0 = Wc[h * 2], w1 = Wc[h * 2 + 1];
    out[h] = t.s * (w0 * t.c + w1 * t.sn);
    out[Hd + h] = t.s * (-w0 * t.sn + w1 * t.c);
    out[2 * Hd + h] = fmaf(w0, t.dx, fmaf(w1, t.dy, v));
  } else {
    float w0 = Wc[h];
    out[h] = w0;
    out[Hd + h] = 0.f;
    out[2 * Hd + h] = fmaf(w0, t.dx, v);
  }
}

}  // namespace pvb
