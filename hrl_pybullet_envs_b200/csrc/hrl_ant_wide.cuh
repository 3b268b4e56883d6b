// hrl_ant_wide.cuh - the Ant sub-step in the WIDE mappings: 4 * SUB lanes per env (SUB = 2: 8 lanes, 4 envs per warp;
// SUB = 4: 16 lanes, 2 envs per warp) instead of the 4 lanes per env of hrl_ant.cuh.
//
// Why: at the metric's batch size (4096 envs) the 4-lane mapping gives 512 warps for 592 warp schedulers, so the step
// time is the latency of ONE warp's instruction stream.  More lanes per env = more warps per scheduler (1.7 / 3.5) to
// hide that latency, provided the extra lanes take work off the stream:
//   * lane (k, sub) = leg k (lane & 3), sub-lane sub (bits above): the per-leg smooth dynamics is replicated in the SUB
//     sub-lanes of a leg (same registers, same values);
//   * contact detection: the 4 sphere slots of a leg (torso, tip, ankle, hip) are split over the sub-lanes, the
//     candidates are compacted into the leg's list in Bullet's order with a prefix sum over the sub-lanes;
//   * row build: contact c of leg k is whitened by sub-lane c % SUB, joint-limit row j by sub-lane j - a lane builds
//     <= 3 * (4 / SUB) + 1 rows instead of <= 14;
//   * solver: a row only touches the base (6) and ONE leg (2), so rows are stored 8 wide and the transformed velocity is
//     kept as [base part, replicated in every lane | own leg's part]: the lanes of the row's leg hold its true residual,
//     the impulse delta reaches the other lanes with one shuffle.  Visit order, clamps and the friction cone are those
//     of hrl_ant.cuh (Bullet's), so the same parity tests gate both mappings.
#pragma once
#include "hrl_ant.cuh"

// rows of one env: same visit-order positions as hrl_ant.cuh (limits 0-7 | friction pairs 8+2c, 9+2c | zero rows 40, 41 |
// normals descending 57-c), 3 float4 each: (z0 z1 z2 z3) (z4 z5 y0 y1) (1/diag, rhs/diag, leg, -)
#define HRLW_ROW_F4 3
#define HRLW_ENV_F4 (HRL_ROWS_ENV * HRLW_ROW_F4 + 1)
template <int SUB>
struct WideMap {
  static constexpr int LPE = 4 * SUB;      // lanes per env
  static constexpr int EPW = 32 / LPE;     // envs per warp
  static constexpr int SPL = 4 / SUB;      // sphere slots per lane
  static constexpr int IPL = 16 / LPE;     // food / poison items per lane
  static constexpr int ROWS_FLOATS = EPW * HRLW_ENV_F4 * 4;
  static constexpr int CL_FLOATS = EPW * HRL_CL_STRIDE;
  static constexpr int LAM_FLOATS = (CL_FLOATS + EPW * HRL_LAML_STRIDE + 3) / 4 * 4;
  static constexpr int CAND_FLOATS = HRL_MAXC * HRL_CAND_F * 4 * EPW;  // [c][field][env slot * 4 + leg]
  static constexpr int ITEM_FLOATS = EPW * 32 + EPW * 4;
  static constexpr int SMEM_FLOATS = ROWS_FLOATS + LAM_FLOATS + CAND_FLOATS + ITEM_FLOATS;
};

// sum over ALL lanes of an env (gsum of hrl_math.cuh sums over the 4 legs of one sub-lane group)
template <int LPE>
__device__ __forceinline__ float esum(float v) {
#pragma unroll
  for (int o = 1; o < LPE; o <<= 1) v += __shfl_xor_sync(HRL_FULL_MASK, v, o);
  return v;
}

// candidate list of leg `gl` (= env slot * 4 + leg), GS = 4 * EPW lists per warp
#define CANDW(c, f) cands[((c) * HRL_CAND_F + (f)) * GS + gl]

// Count a candidate; store it when its slot (base + n) lies inside the leg's list.  A counting pass calls this with a
// very negative base (nothing is stored), the emitting pass with the sub-lane's offset in the leg's list.
template <int GS>
__device__ __forceinline__ int add_cand_w(float* __restrict__ cands, int gl, int base, int n, V3 crel, float r, V3 nrm, float dist, float body) {
  const int slot = base + n;
  if (slot >= 0 && slot < HRL_MAXC) {
    const V3 Prel = crel - r * nrm;
    CANDW(slot, 0) = Prel.x; CANDW(slot, 1) = Prel.y; CANDW(slot, 2) = Prel.z;
    CANDW(slot, 3) = nrm.x; CANDW(slot, 4) = nrm.y; CANDW(slot, 5) = nrm.z;
    CANDW(slot, 6) = dist; CANDW(slot, 7) = body;
  }
  return n + 1;
}
template <int GS>
__device__ __noinline__ int sphere_vs_walls_w(V3 c, V3 crel, float r, float body, float wx, float wy, float margin,
                                              float* __restrict__ cands, int gl, int base, int n) {
  float d;
  d = wx - c.x - r; if (d < margin) n = add_cand_w<GS>(cands, gl, base, n, crel, r, mk(-1.f, 0.f, 0.f), d, body);
  d = c.x + wx - r; if (d < margin) n = add_cand_w<GS>(cands, gl, base, n, crel, r, mk(1.f, 0.f, 0.f), d, body);
  d = wy - c.y - r; if (d < margin) n = add_cand_w<GS>(cands, gl, base, n, crel, r, mk(0.f, -1.f, 0.f), d, body);
  d = c.y + wy - r; if (d < margin) n = add_cand_w<GS>(cands, gl, base, n, crel, r, mk(0.f, 1.f, 0.f), d, body);
  return n;
}
// Sphere against an axis-aligned box; bit 8 of the result = the sphere is within the margin (see sphere_vs_aabb_inl).
template <int GS>
__device__ __noinline__ int sphere_vs_aabb_w(V3 c, V3 crel, float r, float body, float lox, float loy, float loz, float hix, float hiy,
                                             float hiz, float margin, float* __restrict__ cands, int gl, int base, int n) {
  const float cc[3] = {c.x, c.y, c.z}, lo[3] = {lox, loy, loz}, hi[3] = {hix, hiy, hiz};
  float qq[3];
  bool inside = true;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float xx = cc[i];
    if (xx < lo[i]) { xx = lo[i]; inside = false; }
    if (xx > hi[i]) { xx = hi[i]; inside = false; }
    qq[i] = xx;
  }
  V3 nrm; float dist;
  if (!inside) {
    const V3 d = mk(cc[0] - qq[0], cc[1] - qq[1], cc[2] - qq[2]);
    const float len = sqrtf(dot(d, d));
    nrm = (1.0f / len) * d; dist = len - r;
  } else {
    float best = 1e30f; int bi = 0; float bs = 1.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float dl = cc[i] - lo[i], dh = hi[i] - cc[i];
      if (dl < best) { best = dl; bi = i; bs = -1.f; }
      if (dh < best) { best = dh; bi = i; bs = 1.f; }
    }
    nrm = mk(bi == 0 ? bs : 0.f, bi == 1 ? bs : 0.f, bi == 2 ? bs : 0.f);
    dist = -best - r;
  }
  if (dist < margin) n = add_cand_w<GS>(cands, gl, base, n, crel, r, nrm, dist, body) | 0x100;
  return n;
}
// Maze box: capsule cylinders of one leg against the box's vertical edges (see capsules_vs_box_edges of hrl_ant.cuh).
template <int GS>
__device__ __noinline__ int capsules_vs_box_edges_w(V3 O, V3 rh, V3 r_ank, V3 r_tip, float lox, float loy, float loz, float hix, float hiy,
                                                    float hiz, float margin, float* __restrict__ cands, int gl, int base, int n) {
#pragma unroll 1
  for (int cap = 0; cap < 3; cap++) {
    const V3 A = cap == 0 ? r_ank : (cap == 1 ? rh : mk(0.f, 0.f, 0.f));
    const V3 B = cap == 0 ? r_tip : (cap == 1 ? r_ank : rh);
    const V3 d = B - A;
    const float L2 = d.x * d.x + d.y * d.y;
    if (!(L2 > 1e-12f)) continue;
    const float iL2 = 1.0f / L2;
#pragma unroll 1
    for (int corner = 0; corner < 4; corner++) {
      const float cx = ((corner & 1) ? hix : lox) - O.x, sgx = (corner & 1) ? 1.f : -1.f;
      const float cy = ((corner & 2) ? hiy : loy) - O.y, sgy = (corner & 2) ? 1.f : -1.f;
      const float t = ((cx - A.x) * d.x + (cy - A.y) * d.y) * iL2;
      if (!(t > 0.f && t < 1.f)) continue;
      const V3 Q = A + t * d;
      const float qz = O.z + Q.z;
      if (qz < loz || qz > hiz) continue;
      const float ex = Q.x - cx, ey = Q.y - cy;
      if (ex * sgx < 0.f || ey * sgy < 0.f) continue;
      const float e2 = ex * ex + ey * ey;
      if (!(e2 > 0.f)) continue;
      const float el = sqrtf(e2), dist = el - ant::R_CAPS;
      if (dist < margin) {
        const float nx = ex / el, ny = ey / el;
        // (the stored contact point is Q - r n: pass the sphere-style arguments crel = Q, r = R_CAPS)
        n = add_cand_w<GS>(cands, gl, base, n, Q, ant::R_CAPS, mk(nx, ny, 0.f), dist, (float)(2 - cap));
      }
    }
  }
  return n;
}

// The sphere slots [si0, si1) of leg k (0 torso - leg 0 only -, 1 foot tip, 2 ankle, 3 hip) against ground / walls /
// maze box / cubes, in the oracle's order.  Returns the number of candidates; `feet` is set when a foot-link sphere
// (tip or ankle) is within the margin of the ground.
template <int GS, bool ITEMS>
__device__ __forceinline__ int leg_spheres_w(const AntLane& s, const SubstepParams& P, V3 rh, V3 r_ank, V3 r_tip, int k, int si0, int si1,
                                             int base, float* __restrict__ cands, int gl, unsigned imask, const float* __restrict__ ixy,
                                             unsigned long long* itouch, bool count_touch, int& feet) {
  int n = 0;
#pragma unroll 1
  for (int si = si0; si < si1; si++) {
    const V3 crel = si == 0 ? mk(0.f, 0.f, 0.f) : (si == 1 ? r_tip : (si == 2 ? r_ank : rh));
    const float r = si == 0 ? (k == 0 ? ant::R_TORSO : -1e6f) : ant::R_CAPS;  // radius -1e6: every distance is huge
    const float body = si == 1 ? 2.f : (si == 2 ? 1.f : 0.f);
    const V3 c = s.O + crel;
    const float dg = c.z - P.gz - r;
    if (dg < P.margin) {
      if (si == 1 || si == 2) feet = 1;
      n = add_cand_w<GS>(cands, gl, base, n, crel, r, mk(0.f, 0.f, 1.f), dg, body);
    }
    if (P.has_walls && fminf(P.wx - fabsf(c.x), P.wy - fabsf(c.y)) - r < P.margin)
      n = sphere_vs_walls_w<GS>(c, crel, r, body, P.wx, P.wy, P.margin, cands, gl, base, n);
    if (!ITEMS && P.has_box) {
      const float reach = r + P.margin;
      const float ax = fmaxf(fmaxf(P.blo[0] - c.x, c.x - P.bhi[0]), 0.f), ay = fmaxf(fmaxf(P.blo[1] - c.y, c.y - P.bhi[1]), 0.f),
                  az = fmaxf(fmaxf(P.blo[2] - c.z, c.z - P.bhi[2]), 0.f);
      if (reach > 0.f && ax * ax + ay * ay + az * az < reach * reach * 1.00001f)
        n = sphere_vs_aabb_w<GS>(c, crel, r, body, P.blo[0], P.blo[1], P.blo[2], P.bhi[0], P.bhi[1], P.bhi[2], P.margin, cands, gl, base, n) & 0xff;
    }
    if (ITEMS) {
      for (unsigned m = imask; m; m &= m - 1) {
        const int gi = __ffs(m) - 1;
        const float bx = ixy[2 * gi], by = ixy[2 * gi + 1], hh = P.item_half, rr = hh + r + P.margin;
        if (fabsf(c.x - bx) > rr || fabsf(c.y - by) > rr) continue;
        n = sphere_vs_aabb_w<GS>(c, crel, r, body + 4.f, bx - hh, by - hh, P.item_z - hh, bx + hh, by + hh, P.item_z + hh, P.margin, cands, gl,
                                 base, n);
        if (count_touch && (n & 0x100)) HRL_TOUCH_ADD(itouch, gi);
        n &= 0xff;
      }
    }
  }
  return n;
}

// Smooth dynamics of leg k (same arithmetic as the block of that name in hrl_ant.cuh::ant_substep): bias forces, leg
// elimination, base Schur complement S = L L^T, unconstrained velocity update.  Replicated in the sub-lanes of a leg;
// gsum() runs over the 4 legs of one sub-lane group.
__device__ __forceinline__ void leg_dynamics_w(const AntLane& s, const SubstepParams& P, const LegKin& K, float tau1, float tau2,
                                               LegDyn& D, float ub[6], float& u1, float& u2) {
  const V3 a1 = K.ez, a2 = K.a2;
  const V3 r_ac = K.rh + K.r1, rf1 = 2.f * K.r1 + K.r2, r_fc = K.rh + rf1;
  const V3 w = s.w, v = s.v;
  const V3 w1 = w + s.qd1 * a1, w2 = w1 + s.qd2 * a2;
  const V3 wxrh = cross(w, K.rh);
  const V3 a_h = cross(w, wxrh);
  const V3 al1 = s.qd1 * cross(w, a1);
  const V3 w1xr1 = cross(w1, K.r1);
  const V3 t1 = cross(al1, K.r1) + cross(w1, w1xr1);
  const V3 a_ac = a_h + t1, a_ank = a_h + 2.f * t1;
  const V3 al2 = al1 + s.qd2 * cross(w1, a2);
  const V3 w2xr2 = cross(w2, K.r2);
  const V3 a_fc = a_ank + cross(al2, K.r2) + cross(w2, w2xr2);
  const V3 v_h = v + wxrh, v_ac = v_h + w1xr1, v_ank = v_h + 2.f * w1xr1, v_fc = v_ank + w2xr2;
  const V3 v_lc = v + 0.5f * wxrh;
  const V3 gz = mk(0.f, 0.f, P.g);
  const V3 Iw1 = axisym(ant::IX_SHORT, ant::IZ_SHORT, K.ez, w1);
  const V3 Iw2 = axisym(ant::IX_LONG, ant::IZ_LONG, K.zf, w2);
  const V3 Iwl = axisym(ant::IX_SHORT, ant::IZ_SHORT, K.ez, w);
  const V3 F_a = ant::M_SHORT * (a_ac + gz) + (P.kl * ant::M_SHORT * (1.f + norm(v_ac))) * v_ac;
  const V3 N_a = axisym(ant::IX_SHORT, ant::IZ_SHORT, K.ez, al1) + cross(w1, Iw1) + (P.ka * (1.f + norm(w1))) * Iw1;
  const V3 F_f = ant::M_LONG * (a_fc + gz) + (P.kl * ant::M_LONG * (1.f + norm(v_fc))) * v_fc;
  const V3 N_f = axisym(ant::IX_LONG, ant::IZ_LONG, K.zf, al2) + cross(w2, Iw2) + (P.ka * (1.f + norm(w2))) * Iw2;
  const V3 F_l = (P.kl * ant::M_SHORT * (1.f + norm(v_lc))) * v_lc;
  const V3 N_l = (P.ka * (1.f + norm(w))) * Iwl;
  const float cb2 = dot(a2, N_f + cross(K.r2, F_f));
  const float cb1 = dot(a1, N_a + cross(K.r1, F_a) + N_f + cross(rf1, F_f));
  V3 cF = F_a + F_f + F_l;
  V3 cT = N_a + cross(r_ac, F_a) + N_f + cross(r_fc, F_f) + N_l + cross(0.5f * K.rh, F_l);

  const V3 lam2 = cross(a2, K.r2), lam1a = cross(a1, K.r1), lam1f = cross(a1, rf1);
  const V3 If_a1 = axisym(ant::IX_LONG, ant::IZ_LONG, K.zf, a1);
  const V3 If_a2 = ant::IX_LONG * a2;
  const float M22 = ant::M_LONG * dot(lam2, lam2) + ant::IX_LONG;
  const float M12 = ant::M_LONG * dot(lam1f, lam2) + dot(a1, If_a2);
  const float M11 = ant::M_SHORT * dot(lam1a, lam1a) + ant::IZ_SHORT + ant::M_LONG * dot(lam1f, lam1f) + dot(a1, If_a1);
  const float idet = rcp_ftz(M11 * M22 - M12 * M12);
  D.mi11 = M22 * idet; D.mi22 = M11 * idet; D.mi12 = -M12 * idet;
  D.il11 = rsqrt_ftz(M11); D.l21 = M12 * D.il11; D.il22 = rsqrt_ftz(M22 - D.l21 * D.l21);
  const V3 G2f = ant::M_LONG * lam2, G2t = cross(r_fc, G2f) + If_a2;
  const V3 G1fa = ant::M_SHORT * lam1a, G1ff = ant::M_LONG * lam1f;
  const V3 G1f = G1fa + G1ff, G1t = cross(r_ac, G1fa) + ant::IZ_SHORT * a1 + cross(r_fc, G1ff) + If_a1;
  D.G1[0] = G1t.x; D.G1[1] = G1t.y; D.G1[2] = G1t.z; D.G1[3] = G1f.x; D.G1[4] = G1f.y; D.G1[5] = G1f.z;
  D.G2[0] = G2t.x; D.G2[1] = G2t.y; D.G2[2] = G2t.z; D.G2[3] = G2f.x; D.G2[4] = G2f.y; D.G2[5] = G2f.z;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    D.K0[i] = D.G1[i] * D.mi11 + D.G2[i] * D.mi12;
    D.K1[i] = D.G1[i] * D.mi12 + D.G2[i] * D.mi22;
  }
  float S[21];
  {
    const float ma = ant::M_SHORT, mf = ant::M_LONG;
    const float da2 = dot(r_ac, r_ac), df2 = dot(r_fc, r_fc);
    const float dza = ant::IZ_SHORT - ant::IX_SHORT, dzf = ant::IZ_LONG - ant::IX_LONG;
    const float diag0 = ant::IX_SHORT + ant::IX_LONG + ma * da2 + mf * df2;
    const float ra[3] = {r_ac.x, r_ac.y, r_ac.z}, rf[3] = {r_fc.x, r_fc.y, r_fc.z};
    const float za[3] = {K.ez.x, K.ez.y, K.ez.z}, zf[3] = {K.zf.x, K.zf.y, K.zf.z};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i; j < 3; j++)
        S[sym6(i, j)] = (i == j ? diag0 : 0.f) + dza * za[i] * za[j] + dzf * zf[i] * zf[j] - ma * ra[i] * ra[j] - mf * rf[i] * rf[j];
    const float mdx = ma * ra[0] + mf * rf[0], mdy = ma * ra[1] + mf * rf[1], mdz = ma * ra[2] + mf * rf[2];
    S[sym6(0, 3)] = 0.f;  S[sym6(0, 4)] = -mdz; S[sym6(0, 5)] = mdy;
    S[sym6(1, 3)] = mdz;  S[sym6(1, 4)] = 0.f;  S[sym6(1, 5)] = -mdx;
    S[sym6(2, 3)] = -mdy; S[sym6(2, 4)] = mdx;  S[sym6(2, 5)] = 0.f;
    S[sym6(3, 3)] = ma + mf; S[sym6(3, 4)] = 0.f; S[sym6(3, 5)] = 0.f;
    S[sym6(4, 4)] = ma + mf; S[sym6(4, 5)] = 0.f; S[sym6(5, 5)] = ma + mf;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = i; j < 6; j++) S[sym6(i, j)] -= D.K0[i] * D.G1[j] + D.K1[i] * D.G2[j];
  }
#pragma unroll
  for (int i = 0; i < 21; i++) S[i] = gsum(S[i]);
  {
    const float dzc = ant::IZ_COMP - ant::IX_COMP;
    const float za[3] = {K.ez.x, K.ez.y, K.ez.z};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i; j < 3; j++) S[sym6(i, j)] += (i == j ? ant::IX_COMP : 0.f) + dzc * za[i] * za[j];
    S[sym6(3, 3)] += ant::M_COMP; S[sym6(4, 4)] += ant::M_COMP; S[sym6(5, 5)] += ant::M_COMP;
  }
  {
    float L[6][6];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) L[i][j] = S[sym6(j, i)];
    float rd[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      float d = L[j][j];
#pragma unroll
      for (int kk = 0; kk < j; kk++) d = fmaf(-L[j][kk], L[j][kk], d);
      rd[j] = rsqrt_ftz(d);
      L[j][j] = d * rd[j];
#pragma unroll
      for (int i = j + 1; i < 6; i++) {
        float a = L[i][j];
#pragma unroll
        for (int kk = 0; kk < j; kk++) a = fmaf(-L[i][kk], L[j][kk], a);
        L[i][j] = a * rd[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 6; j++) {
      D.Li[j * (j + 1) / 2 + j] = rd[j];
#pragma unroll
      for (int i = j + 1; i < 6; i++) {
        float a = 0.f;
#pragma unroll
        for (int kk = j; kk < i; kk++) a = fmaf(L[i][kk], D.Li[kk * (kk + 1) / 2 + j], a);
        D.Li[i * (i + 1) / 2 + j] = -a * rd[i];
      }
    }
  }
  const float f1 = tau1 - cb1, f2 = tau2 - cb2;
  const float y1 = D.mi11 * f1 + D.mi12 * f2, y2 = D.mi12 * f1 + D.mi22 * f2;
  float fb[6] = {-cT.x, -cT.y, -cT.z, -cF.x, -cF.y, -cF.z};
#pragma unroll
  for (int i = 0; i < 6; i++) fb[i] = gsum(fb[i] - (D.G1[i] * y1 + D.G2[i] * y2));
  {
    const V3 IwT = axisym(ant::IX_COMP, ant::IZ_COMP, K.ez, w);
    const V3 tT = cross(w, IwT) + (P.ka * (1.f + norm(w)) * ant::I_TORSO) * w;
    const V3 fT = ant::M_COMP * gz + (P.kl * ant::M_TORSO * (1.f + norm(v))) * v;
    fb[0] -= tT.x; fb[1] -= tT.y; fb[2] -= tT.z; fb[3] -= fT.x; fb[4] -= fT.y; fb[5] -= fT.z;
  }
  float ud[6];
  sinv_mul(D.Li, fb, ud);
  float qdd1 = y1, qdd2 = y2;
#pragma unroll
  for (int i = 0; i < 6; i++) { qdd1 = fmaf(-D.K0[i], ud[i], qdd1); qdd2 = fmaf(-D.K1[i], ud[i], qdd2); }
  const float wv[6] = {w.x, w.y, w.z, v.x, v.y, v.z};
#pragma unroll
  for (int i = 0; i < 6; i++) ub[i] = clampf(fmaf(P.h, ud[i], wv[i]), P.vmax);
  u1 = clampf(fmaf(P.h, qdd1, s.qd1), P.vmax);
  u2 = clampf(fmaf(P.h, qdd2, s.qd2), P.vmax);
}

// Whiten one constraint row of leg k (see emit_row of hrl_ant.cuh) and store it 8 wide at visit position `pos`.
__device__ __forceinline__ void emit_row_w(float4* __restrict__ rb, int pos, int k, const LegDyn& D, const float JB[6], float j1, float j2,
                                           const float ub[6], float u1, float u2, float pen, float erp, float inv_h, bool positional) {
  HRL_CHECK(pos >= 0 && pos < HRL_ROWS_ENV && pos != HRL_ROW_ZERO && pos != HRL_ROW_ZERO + 1);
  float Jt[6], z[6];
#pragma unroll
  for (int i = 0; i < 6; i++) Jt[i] = JB[i] - (D.K0[i] * j1 + D.K1[i] * j2);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j <= i; j++) a = fmaf(D.Li[i * (i + 1) / 2 + j], Jt[j], a);
    z[i] = a;
  }
  const float y0 = j1 * D.il11, y1 = (j2 - D.l21 * y0) * D.il22;
  float diag = y0 * y0 + y1 * y1, rel = j1 * u1 + j2 * u2;
#pragma unroll
  for (int i = 0; i < 6; i++) { diag = fmaf(z[i], z[i], diag); rel = fmaf(JB[i], ub[i], rel); }
  const float dinv = rcp_ftz(diag);
  float posErr = 0.f, velErr = -rel;
  if (positional) {
    if (pen > 0.f) velErr -= pen * inv_h;
    else posErr = -pen * erp * inv_h;
  }
  float4* r = rb + pos * HRLW_ROW_F4;
  r[0] = make_float4(z[0], z[1], z[2], z[3]);
  r[1] = make_float4(z[4], z[5], y0, y1);
  r[2] = make_float4(dinv, (posErr + velErr) * dinv, __int_as_float(k), 0.f);
}

struct RowW { float2 p[4]; float dinv, rhs; int leg; };  // (z01 z23 z45 y01), 1/diag, rhs/diag, leg the row belongs to
__device__ __forceinline__ RowW ld_row_w(const float4* __restrict__ rb, int r) {
  const float4 q0 = rb[r * HRLW_ROW_F4], q1 = rb[r * HRLW_ROW_F4 + 1], q2 = rb[r * HRLW_ROW_F4 + 2];
  RowW R;
  R.p[0] = make_float2(q0.x, q0.y); R.p[1] = make_float2(q0.z, q0.w); R.p[2] = make_float2(q1.x, q1.y); R.p[3] = make_float2(q1.z, q1.w);
  R.dinv = q2.x; R.rhs = q2.y; R.leg = __float_as_int(q2.z);
  return R;
}
// residual of a row against the transformed velocity [dvb (3 pairs) | g of THIS lane's leg]: true in the lanes of the row's leg
__device__ __forceinline__ float dot8(const float2 a[4], const float2 dvb[3], float2 g) {
  float2 s0 = mul2(a[0], dvb[0]), s1 = mul2(a[1], dvb[1]);
  s0 = fma2(a[2], dvb[2], s0); s1 = fma2(a[3], g, s1);
  const float2 t = add2(s0, s1);
  return t.x + t.y;
}
__device__ __forceinline__ void axpy8(float2 dvb[3], float2& g, const float2 a[4], float dl, bool mine) {
  const float2 ss = make_float2(dl, dl), sg = mine ? ss : make_float2(0.f, 0.f);
  dvb[0] = fma2(a[0], ss, dvb[0]); dvb[1] = fma2(a[1], ss, dvb[1]); dvb[2] = fma2(a[2], ss, dvb[2]);
  g = fma2(a[3], sg, g);
}
// One visit of a single row.  Every lane evaluates it; the lanes of the row's leg hold the true residual, their impulse
// delta is what everybody applies (one shuffle from sub-lane 0 of that leg).  `ebase` = first lane of this env.
template <bool HAS_HI>
__device__ __forceinline__ void single_visit_w(float* __restrict__ lp, float2 dvb[3], float2& g, const RowW& R, float lam, float hi, int k, int ebase) {
  float nl = fmaxf(fmaf(-dot8(R.p, dvb, g), R.dinv, lam + R.rhs), 0.f);
  if (HAS_HI) nl = fminf(nl, hi);
  const bool mine = (k == R.leg);
  if (mine) *lp = nl;
  const float dl = __shfl_sync(HRL_FULL_MASK, nl - lam, ebase | R.leg);
  axpy8(dvb, g, R.p, dl, mine);
}
__device__ __forceinline__ void pair_visit_w(float4* __restrict__ cp, float2 dvb[3], float2& g, const RowW& A, const RowW& B, const float4 c, int k, int ebase) {
  float na = fmaf(-dot8(A.p, dvb, g), A.dinv, c.x + A.rhs);
  float nb = fmaf(-dot8(B.p, dvb, g), B.dinv, c.y + B.rhs);
  const float lim = c.w * c.z, len2 = fmaf(na, na, nb * nb);
  const float sc = (len2 > lim * lim) ? lim * rsqrt_ftz(fmaxf(len2, 1e-30f)) : 1.f;
  const bool on = c.z > 0.f;
  na = on ? na * sc : c.x; nb = on ? nb * sc : c.y;
  const bool mine = (k == A.leg);
  if (mine) *reinterpret_cast<float2*>(cp) = make_float2(na, nb);
  const float da = __shfl_sync(HRL_FULL_MASK, na - c.x, ebase | A.leg), db = __shfl_sync(HRL_FULL_MASK, nb - c.y, ebase | A.leg);
  axpy8(dvb, g, A.p, da, mine); axpy8(dvb, g, B.p, db, mine);
}

// One internal step of h = dt/substeps in the wide mapping.  rows / cands / iscr: this WARP's shared memory
// (WideMap<SUB> sizes).  l = lane within the env, k = l & 3 (leg), sub = l >> 2, es = env slot in the warp.
template <int SUB, bool ITEMS>
__device__ __forceinline__ void ant_substep_w(AntLane& s, const SubstepParams& P, const LegConst& lc, float tau1, float tau2,
                                              float* __restrict__ rows, float* __restrict__ cands, int lane, int l, int k, int sub, int es,
                                              int& feet_ground, int& stat_contacts, int& stat_limits, const float* it_x = nullptr,
                                              const float* it_y = nullptr, float* iscr = nullptr, bool count_touch = false, float* dbg = nullptr) {
  typedef WideMap<SUB> M;
  constexpr int LPE = M::LPE, GS = 4 * M::EPW;
  const int ebase = lane & ~(LPE - 1), gl = es * 4 + k;
  const LegKin K = leg_fk(s, lc);
  const V3 a1 = K.ez, a2 = K.a2;
  const V3 r_ank = K.rh + 2.f * K.r1;
  const V3 r_tip = r_ank + 2.f * K.r2;

  // ---------------- contacts: the leg's sphere slots split over its sub-lanes ----------------
  int nC;  // contacts of this leg (<= HRL_MAXC), known to all its sub-lanes
  {
    unsigned imask = 0;
    float* ixy = nullptr;
    unsigned long long* itouch = nullptr;
    if (ITEMS && P.item_contacts) {
      ixy = iscr + es * 32; itouch = reinterpret_cast<unsigned long long*>(iscr + M::EPW * 32) + 2 * es;
      const float reach = 1.1314f + 1.4143f * (P.item_half + ant::R_CAPS + P.margin);
      if (sub * 4 + k == l) {  // (always true; keeps the compiler from hoisting the loads above the branch)
#pragma unroll
        for (int i = 0; i < M::IPL; i++) {
          const int gi = M::IPL * l + i;
          const float dx = it_x[i] - s.O.x, dy = it_y[i] - s.O.y;
          if (gi < P.n_items && dx * dx + dy * dy < reach * reach) imask |= 1u << gi;
          ixy[2 * gi] = it_x[i]; ixy[2 * gi + 1] = it_y[i];
        }
      }
      if (count_touch && l == 0) { itouch[0] = 0ull; itouch[1] = 0ull; }
#pragma unroll
      for (int o = 1; o < LPE; o <<= 1) imask |= __shfl_xor_sync(HRL_FULL_MASK, imask, o);
      __syncwarp();
    }
    const int si0 = sub * M::SPL, si1 = si0 + M::SPL;
    int feet = 0;
    // pass 1 counts (nothing stored), the prefix over the sub-lanes gives this lane's offset in the leg's list, pass 2 stores
    const int n_mine = leg_spheres_w<GS, ITEMS>(s, P, K.rh, r_ank, r_tip, k, si0, si1, -1024, cands, gl, imask, ixy, itouch, false, feet);
    int incl = n_mine;
#pragma unroll
    for (int d = 4; d < LPE; d <<= 1) {
      const int t = __shfl_up_sync(HRL_FULL_MASK, incl, d, LPE);
      if (l >= d) incl += t;
    }
    int total = __shfl_sync(HRL_FULL_MASK, incl, ebase | (LPE - 4) | k);  // last sub-lane of this leg
#pragma unroll
    for (int o = 4; o < LPE; o <<= 1) feet |= __shfl_xor_sync(HRL_FULL_MASK, feet, o);
    feet_ground = feet;
    // warp-uniform on purpose.  Guarded per lane (`if (n_mine > 0)`) this pass miscomputed on the B200 whenever only some
    // lanes took the branch around the out-of-line cube tests (measured with tools/debug_wide.py: identical candidate
    // lists, different solves); with every lane inside, the wide and the 4-lane mapping agree to rounding.
    if (__any_sync(HRL_FULL_MASK, n_mine > 0)) {
      int dummy = 0;
      leg_spheres_w<GS, ITEMS>(s, P, K.rh, r_ank, r_tip, k, si0, si1, incl - n_mine, cands, gl, imask, ixy, itouch, count_touch, dummy);
    }
    if (ITEMS) {  // capsule cylinders vs the cubes within reach: after the spheres, by sub-lane 0 of the leg
      int ne = 0;
      if (sub == 0 && imask)
        ne = capsules_vs_cubes(s.O, K.rh, r_ank, r_tip, imask, ixy, P.item_half, P.item_z, P.margin, cands + gl, GS, total, itouch, count_touch);
      total += __shfl_sync(HRL_FULL_MASK, ne, ebase | k);
    }
    if (!ITEMS && P.has_box) {  // capsule cylinders vs the box's vertical edges: after the spheres, by sub-lane 0 of the leg
      const float reach = 1.1314f + ant::R_CAPS + P.margin;
      const float nx = fmaxf(fmaxf(P.blo[0] - s.O.x, s.O.x - P.bhi[0]), 0.f), ny = fmaxf(fmaxf(P.blo[1] - s.O.y, s.O.y - P.bhi[1]), 0.f);
      int ne = 0;
      if (sub == 0 && nx * nx + ny * ny < reach * reach)
        ne = capsules_vs_box_edges_w<GS>(s.O, K.rh, r_ank, r_tip, P.blo[0], P.blo[1], P.blo[2], P.bhi[0], P.bhi[1], P.bhi[2], P.margin, cands, gl,
                                         total, 0);
      total += __shfl_sync(HRL_FULL_MASK, ne, ebase | k);  // (outside the cull: every lane of the warp takes part in the shuffle)
    }
    nC = min(total, HRL_MAXC);
    __syncwarp();
  }

  if (dbg && sub == 0) {  // debugging builds (HRL_DEBUG_CONTACTS): this leg's candidate list
    dbg[0] = (float)nC;
    for (int c = 0; c < nC; c++)
      for (int f = 0; f < HRL_CAND_F; f++) dbg[1 + c * HRL_CAND_F + f] = CANDW(c, f);
  }
  // ---------------- smooth dynamics (replicated in the sub-lanes of the leg) ----------------
  LegDyn D;
  float ub[6], u1, u2;
  leg_dynamics_w(s, P, K, tau1, tau2, D, ub, u1, u2);

  // ---------------- constraint rows: counts, visit positions; sub-lane c % SUB whitens contact c ----------------
  const float inv_h = P.inv_h;
  float4* __restrict__ rb = reinterpret_cast<float4*>(rows) + es * HRLW_ENV_F4;
  float4* __restrict__ cl = reinterpret_cast<float4*>(rows + M::ROWS_FLOATS + es * HRL_CL_STRIDE);
  float* __restrict__ lamL = rows + M::ROWS_FLOATS + M::CL_FLOATS + es * HRL_LAML_STRIDE;
  const float pl1 = s.q1 - ant::HIP_LO, ph1 = ant::HIP_HI - s.q1;
  const float pl2 = s.q2 - lc.lo2, ph2 = lc.hi2 - s.q2;
  const bool lim1 = (pl1 <= 0.f) || (ph1 <= 0.f), lim2 = (pl2 <= 0.f) || (ph2 <= 0.f);
  const int nL = (int)lim1 + (int)lim2;
  int incl = nL | (nC << 8);  // inclusive scan over the 4 legs (of this sub-lane group) -> Bullet row order
  {
    int t = __shfl_up_sync(HRL_FULL_MASK, incl, 1, 4);
    if (k >= 1) incl += t;
    t = __shfl_up_sync(HRL_FULL_MASK, incl, 2, 4);
    if (k >= 2) incl += t;
  }
  const int tot = __shfl_sync(HRL_FULL_MASK, incl, 3, 4);
  const int offL = (incl & 0xff) - nL, offC = (incl >> 8) - nC, NL = tot & 0xff, NC = tot >> 8;
  if (sub < 2 && (sub ? lim2 : lim1)) {  // joint-limit row of joint `sub`
    const float zero6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float pl = sub ? pl2 : pl1, ph = sub ? ph2 : ph1;
    const bool lo = pl <= 0.f;
    const float sg = lo ? 1.f : -1.f, pen = lo ? pl : ph;
    const int pos = offL + (sub ? (int)lim1 : 0);
    lamL[pos] = 0.f;
    emit_row_w(rb, pos, k, D, zero6, sub ? 0.f : sg, sub ? sg : 0.f, ub, u1, u2, pen, P.erp_l, inv_h, true);
  }
#pragma unroll 1
  for (int c = sub; c < nC; c += SUB) {
    const V3 Pr = mk(CANDW(c, 0), CANDW(c, 1), CANDW(c, 2));
    const V3 n = mk(CANDW(c, 3), CANDW(c, 4), CANDW(c, 5));
    const float dist = CANDW(c, 6), bodyc = CANDW(c, 7);
    const bool cube = ITEMS && bodyc >= 4.f;
    const float body = cube ? bodyc - 4.f : bodyc;
    V3 t1, t2;
    plane_space(n, t1, t2);
    const V3 Ph = Pr - K.rh, Pa = Pr - r_ank;
    const int ci = offC + c;
    cl[ci] = make_float4(0.f, 0.f, 0.f, cube ? P.mu_item : P.mu);
#pragma unroll 1
    for (int di = 0; di < 3; di++) {
      const V3 d = di == 0 ? n : (di == 1 ? t1 : t2);
      const V3 jt = cross(Pr, d);
      const float JB[6] = {jt.x, jt.y, jt.z, d.x, d.y, d.z};
      const float j1 = body >= 1.f ? dot(a1, cross(Ph, d)) : 0.f;
      const float j2 = body >= 2.f ? dot(a2, cross(Pa, d)) : 0.f;
      const int pos = di == 0 ? HRL_ROW_NRM_LAST - ci : HRL_ROW_FRI0 + 2 * ci + (di - 1);
      emit_row_w(rb, pos, k, D, JB, j1, j2, ub, u1, u2, dist, P.erp_c, inv_h, di == 0);
    }
  }
  if (l == 0) { stat_contacts += NC; stat_limits += NL; }
  __syncwarp();

  // ---------------- projected Gauss-Seidel, Bullet row order (see hrl_ant.cuh) ----------------
  float2 dvb[3], g = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 3; i++) dvb[i] = make_float2(0.f, 0.f);
  const int maxNL = __reduce_max_sync(HRL_FULL_MASK, NL), maxNC = __reduce_max_sync(HRL_FULL_MASK, NC);
#define HRL_LIM_ROW(t) (((t) < NL) ? lbase + lstep * (t) : HRL_ROW_ZERO)
#define HRL_LIM_LAM(t, r) (lamL + (((t) < NL) ? (r) : 8))
#define HRL_SLOT(t) (((t) < NC) ? (t) : HRL_NSLOT)
  for (int it = 0; it < P.iters; it++) {
    if (maxNL > 0) {
      const int lbase = (it & 1) ? 0 : NL - 1, lstep = (it & 1) ? 1 : -1;
      int r0 = HRL_LIM_ROW(0), r1;
      float *p0 = HRL_LIM_LAM(0, r0), *p1;
      RowW R0 = ld_row_w(rb, r0), R1;
      float l0 = *p0, l1;
      int t = 0;
      for (; t + 1 < maxNL; t += 2) {
        r1 = HRL_LIM_ROW(t + 1); p1 = HRL_LIM_LAM(t + 1, r1); R1 = ld_row_w(rb, r1); l1 = *p1;
        single_visit_w<true>(p0, dvb, g, R0, l0, P.max_imp, k, ebase);
        r0 = HRL_LIM_ROW(t + 2); p0 = HRL_LIM_LAM(t + 2, r0); R0 = ld_row_w(rb, r0); l0 = *p0;
        single_visit_w<true>(p1, dvb, g, R1, l1, P.max_imp, k, ebase);
      }
      if (t < maxNL) single_visit_w<true>(p0, dvb, g, R0, l0, P.max_imp, k, ebase);
    }
    if (maxNC > 0) {
      {
        int c0 = HRL_SLOT(0), c1;
        RowW R0 = ld_row_w(rb, HRL_ROW_NRM_LAST - c0), R1;
        float l0 = cl[c0].z, l1;
        int t = 0;
        for (; t + 1 < maxNC; t += 2) {
          c1 = HRL_SLOT(t + 1); R1 = ld_row_w(rb, HRL_ROW_NRM_LAST - c1); l1 = cl[c1].z;
          single_visit_w<false>(&cl[c0].z, dvb, g, R0, l0, 0.f, k, ebase);
          c0 = HRL_SLOT(t + 2); R0 = ld_row_w(rb, HRL_ROW_NRM_LAST - c0); l0 = cl[c0].z;
          single_visit_w<false>(&cl[c1].z, dvb, g, R1, l1, 0.f, k, ebase);
        }
        if (t < maxNC) single_visit_w<false>(&cl[c0].z, dvb, g, R0, l0, 0.f, k, ebase);
      }
      {
        int c0 = HRL_SLOT(0), c1;
        RowW A0 = ld_row_w(rb, HRL_ROW_FRI0 + 2 * c0), B0 = ld_row_w(rb, HRL_ROW_FRI0 + 2 * c0 + 1), A1, B1;
        float4 q0 = cl[c0], q1;
        int t = 0;
        for (; t + 1 < maxNC; t += 2) {
          c1 = HRL_SLOT(t + 1); A1 = ld_row_w(rb, HRL_ROW_FRI0 + 2 * c1); B1 = ld_row_w(rb, HRL_ROW_FRI0 + 2 * c1 + 1); q1 = cl[c1];
          pair_visit_w(cl + c0, dvb, g, A0, B0, q0, k, ebase);
          c0 = HRL_SLOT(t + 2); A0 = ld_row_w(rb, HRL_ROW_FRI0 + 2 * c0); B0 = ld_row_w(rb, HRL_ROW_FRI0 + 2 * c0 + 1); q0 = cl[c0];
          pair_visit_w(cl + c1, dvb, g, A1, B1, q1, k, ebase);
        }
        if (t < maxNC) pair_visit_w(cl + c0, dvb, g, A0, B0, q0, k, ebase);
      }
    }
  }
#undef HRL_LIM_LAM
#undef HRL_SLOT
#undef HRL_LIM_ROW
  __syncwarp();

  // ---------------- back to physical velocities, clamp, integrate (as hrl_ant.cuh) ----------------
  float dvb6[6] = {dvb[0].x, dvb[0].y, dvb[1].x, dvb[1].y, dvb[2].x, dvb[2].y}, dvp[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {  // dvb = L^-T dvb'
    float a = 0.f;
#pragma unroll
    for (int j = i; j < 6; j++) a = fmaf(D.Li[j * (j + 1) / 2 + i], dvb6[j], a);
    dvp[i] = a;
  }
  const float g2 = g.y * D.il22, g1 = (g.x - D.l21 * g2) * D.il11;  // g = Ll^-T g'
  float dq1 = g1, dq2 = g2;
#pragma unroll
  for (int i = 0; i < 6; i++) { dq1 = fmaf(-D.K0[i], dvp[i], dq1); dq2 = fmaf(-D.K1[i], dvp[i], dq2); }
#pragma unroll
  for (int i = 0; i < 6; i++) ub[i] = clampf(ub[i] + dvp[i], P.vmax);
  u1 = clampf(u1 + dq1, P.vmax); u2 = clampf(u2 + dq2, P.vmax);
  s.w = mk(ub[0], ub[1], ub[2]); s.v = mk(ub[3], ub[4], ub[5]);
  s.qd1 = u1; s.qd2 = u2;
  s.O = s.O + P.h * s.v;
  s.q1 = fmaf(P.h, u1, s.q1); s.q2 = fmaf(P.h, u2, s.q2);
  {
    const float w2 = dot(s.w, s.w), y2 = 0.25f * P.h * P.h * w2;
    float kk, cw;
    if (y2 <= 0.1542126f) {
      kk = 0.5f * P.h * fmaf(y2, fmaf(y2, fmaf(y2, fmaf(y2, 2.7557319e-6f, -1.9841270e-4f), 8.3333333e-3f), -0.16666667f), 1.f);
      cw = fmaf(y2, fmaf(y2, fmaf(y2, fmaf(y2, 2.4801587e-5f, -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.f);
    } else {
      const float2 e = quat_exp_literal(w2, P.h);
      kk = e.x; cw = e.y;
    }
    const float ax = s.w.x * kk, ay = s.w.y * kk, az = s.w.z * kk;
    const float x = s.qx, y = s.qy, z = s.qz, qw = s.qw;
    float nx = cw * x + ax * qw + ay * z - az * y;
    float ny = cw * y + ay * qw + az * x - ax * z;
    float nz = cw * z + az * qw + ax * y - ay * x;
    float nw = cw * qw - ax * x - ay * y - az * z;
    const float inv = rsqrt_ftz(nx * nx + ny * ny + nz * nz + nw * nw);
    s.qx = nx * inv; s.qy = ny * inv; s.qz = nz * inv; s.qw = nw * inv;
  }
}
