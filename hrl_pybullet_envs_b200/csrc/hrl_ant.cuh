// hrl_ant.cuh - one Bullet-style internal sub-step of the Ant, 4 lanes per env (one lane per leg).
//
// What this replaces: scene.global_step() -> p.stepSimulation() for the Ant multibody
// (reference call site envs/gather/ant_gather_env.py:78; arithmetic in pybullet, restated in
// oracle/hrl_oracle.c in a different formulation).
//
// Formulation (DESIGN.md "Kernel 1"): world-frame projected Newton-Euler.  The mass matrix of
// the tree (6 base dofs + 4 legs x 2 hinges) is block-arrow: M = [[Mbb, G],[G^T, blkdiag(Mll_k)]].
// Lane k eliminates its own leg (2x2 Cholesky), contributes its articulated inertia to the base
// Schur complement S (21 unique entries, summed with two xor-shuffles), every lane then holds
// L^-1 of S = L L^T in registers.  A constraint row only touches the base and ONE leg: its owner
// lane whitens it (z = L^-1 Jb~, y = Ll^-1 Jl) and stores it at its position in Bullet's visit
// order.  The projected Gauss-Seidel sweep then runs in the transformed velocity [L^T dvb ; Ll^T g],
// where the same 14-vector is row Jacobian and row response; the 4 lanes of an env evaluate every
// row redundantly from broadcast shared-memory reads (no shuffle on the dependency chain) with
// packed FFMA2 arithmetic.
#pragma once
#include "hrl_math.cuh"
#include "../../include/hrl_b200.h"

#ifndef HRL_MAXC
#define HRL_MAXC 4  // contacts kept per lane (= per contact group, oracle MAX_CONTACT_PER_GROUP)
#endif
// envs per warp (4 lanes each).  8 fills the warp.  With fewer, the spare lane groups shadow an env
// of the same warp (same inputs, same shared-memory slots, identical values; global stores suppressed).
#ifndef HRL_ENVS_PER_WARP
#define HRL_ENVS_PER_WARP 8
#endif
#define HRL_EPW HRL_ENVS_PER_WARP
// Constraint rows of one env live in shared memory in VISIT order (Bullet's row order), whitened:
//   row = a[14] | dinv | rhs  (4 x float4), a = [L^-1 Jb~ (6) ; Ll_k^-1 Jl scattered to leg k (8)]
// so that "J M^-1 J^T" of two rows is a plain dot product and the same vector is both the row
// Jacobian and the row response in the transformed velocity dv' = [L^T dvb ; Ll^T g].
//   rows [0,8)    joint-limit rows (<= 2 per leg), ordered by leg
//   rows [8,40)   friction pairs (2 per contact): contact c at rows 8 + 2c, 9 + 2c (c ordered by leg then candidate)
//   rows 40, 41   all-zero rows: what an idle visit reads (idle friction pair = contact slot 16 = exactly these two)
//   rows [42,58)  contact normals, DEscending: contact c at row 57 - c, so that slot 16 lands on zero row 41
// Every row address is affine in the contact slot, idle included: one select per visit.
#define HRL_NSLOT (4 * HRL_MAXC)                 // contact slots per env; slot HRL_NSLOT is the idle one
#define HRL_ROW_FRI0 8
#define HRL_ROW_ZERO (8 + 2 * HRL_NSLOT)
#define HRL_ROW_NRM_LAST (HRL_ROW_ZERO + 1 + HRL_NSLOT)
#define HRL_ROWS_ENV (HRL_ROW_NRM_LAST + 1)
#define HRL_ENV_F4 (HRL_ROWS_ENV * 4 + 1)        // float4 per env, +1: the env stride maps the 8 envs of a warp to distinct banks
// COMPACT rows (large batches): a row only touches the base and ONE leg, so it is stored as its 8 non-zeros + the leg
// index in 3 float4 - (z0 z1 z2 z3) (z4 z5 y0 y1) (1/diag, rhs/diag, leg, -) - and expanded to the 14-wide form by four
// selects when it is loaded.  7.4 KB less shared memory per warp: 7 instead of 6 CTAs per SM.
#define HRL_ENV_F4_COMPACT (HRL_ROWS_ENV * 3 + 1)
// Impulses: per contact slot one float4 (lambda_t1, lambda_t2, lambda_n, mu) - a friction visit needs all four - and
// per env 9 limit-row impulses (slot 8 = idle).  Strides 68 (= 4 mod 32) / 9 (odd): the 8 envs of a warp hit distinct banks.
#define HRL_CL_STRIDE (4 * (HRL_NSLOT + 1))
#define HRL_LAML_STRIDE 9
#define HRL_CL_FLOATS_PER_WARP (HRL_EPW * HRL_CL_STRIDE)
#define HRL_ROWS_FLOATS_PER_WARP (HRL_EPW * HRL_ENV_F4 * 4)
#define HRL_LAM_FLOATS_PER_WARP ((HRL_CL_FLOATS_PER_WARP + HRL_EPW * HRL_LAML_STRIDE + 3) / 4 * 4)
// contact candidates: [c][field][lane]: P-O (3), n (3), dist, body
#define HRL_CAND_F 8
#define HRL_SMEM_FLOATS_PER_WARP (HRL_ROWS_FLOATS_PER_WARP + HRL_LAM_FLOATS_PER_WARP + HRL_MAXC * HRL_CAND_F * 32)

// ---- Delassus-space sweep (fast path of the 4-lane mapping; DESIGN.md "Round 2 - Delassus-space sweep") ----
// When no env of the warp has more than 4 contacts in this sub-step (warp-uniform test; in the settled regime that is
// every warp: 4 feet on the ground), the rows are NOT stored in visit order but TRANSPOSED into 24 static positions
//   0-7 joint-limit rows (Bullet order), 8 + 4c + (0, 1, 2) normal / friction 1 / friction 2 of contact c (11 + 4c: pad)
// and the sweep runs on the normalised Gram matrix What[i][j] = (a_i . a_j) / (a_j . a_j) (zero diagonal): per row it
// tracks the unclamped target t_j = lambda_j + rhs'_j - (a_j . dv') / (a_j . a_j) in a register, so that a visit is
// clamp(t_i) -> dl -> a few independent FMAs - no 14-term dot and no axpy on the dependency chain (the velocity-space
// visit is dot -> fma -> max -> sub -> axpy, ~60 cycles of latency with one warp per scheduler).  The targets of the
// 8 limit rows live redundantly in all 4 lanes of the env (a limit visit needs no communication); lane c OWNS contact c
// (normal + friction pair: the pair visit finds lambda_n, both targets and mu in its own registers) and broadcasts its
// impulse change with one shuffle.  Same iterates as the velocity-space sweep in exact arithmetic
// (tests/test_delassus_form.py).  The buffers alias the env's row region (floats from its base); the zero rows of the
// velocity-space path are re-zeroed when that path runs.
// MEASURED AND NOT SHIPPED (-DHRL_DELASSUS=1 builds it; tools/gpu_ab_sweep.sh runs the A/B; all GPU parity tests green):
// 51.1 us per step against 43.3 us for the velocity-space sweep at 4096 AntGather envs (a fully redundant variant, all 20
// targets in every lane: 49.8 us).  The Gram matrix costs 13.6 k cycles per step and the shuffles put ~30 cycles of
// latency on every contact visit - with one warp per scheduler nothing hides either (profiles/r2_delassus_*).
#ifndef HRL_DELASSUS
#define HRL_DELASSUS 0
#endif
#ifndef HRL_LOOKAHEAD
#define HRL_LOOKAHEAD 0
#endif
#define HRL_DS_MAXC 4        // contacts per env the fast path holds (more: velocity-space sweep)
#define HRL_DS_P 24          // positions
#define HRL_DS_ROWS 20       // rows of What: limit p -> p, contact c direction d -> 8 + 3 c + d
#define HRL_DS_AT 0          // [14][24] whitened rows, transposed: component m of position p at m * 24 + p
#define HRL_DS_DINV 336      // [24] 1 / (a . a)
#define HRL_DS_RHS 360       // [24] rhs / (a . a)
#define HRL_DS_LEG 384       // [24] leg of the position's row (int)
#define HRL_DS_ZERO_F4 104   // float4 cleared per env and sub-step: everything above, padded to 4 lanes x 26
#define HRL_DS_W 416         // [20][24] What, row r at 416 + 24 r
static_assert(HRL_DS_W + HRL_DS_ROWS * HRL_DS_P <= HRL_ROWS_ENV * 16, "Delassus buffers must fit in the env's row region");
static_assert(HRL_MAXC >= 1 && HRL_DS_MAXC == 4, "one contact per lane");

// food / poison cube colliders (hrl_config.item_contacts): per-warp scratch [EPW][16] (x, y) + [EPW][2] 64-bit words
// holding 16 8-bit contact-point counters (13 spheres + 12 capsule cylinders can touch one cube)
#define HRL_ITEM_SCRATCH_FLOATS (HRL_EPW * 16 * 2 + HRL_EPW * 4)
#define HRL_TOUCH_ADD(itouch, gi) atomicAdd((itouch) + ((gi) >> 3), 1ull << (8 * ((gi) & 7)))
#define HRL_TOUCH_GET(itouch, gi) ((int)(((itouch)[(gi) >> 3] >> (8 * ((gi) & 7))) & 255ull))

struct AntLane {
  // replicated in the 4 lanes of an env
  V3 O;                  // torso origin
  float qx, qy, qz, qw;  // torso orientation
  V3 v, w;               // world-frame linear (of O) / angular velocity
  // private to the lane (leg k)
  float q1, q2, qd1, qd2;  // hip_k, ankle_k
};

struct LegConst {
  float sx, sy;    // ant.xml:15,26,37,48 leg direction signs
  float axl, ayl;  // ankle axis in aux coords (ant.xml:21,32,43,54), normalised
  float lo2, hi2;  // ankle limits
};
__device__ __forceinline__ LegConst leg_const(int k) {
  LegConst c;
  c.sx = (k == 0 || k == 3) ? 1.f : -1.f;
  c.sy = (k < 2) ? 1.f : -1.f;
  c.axl = (k == 0 || k == 2) ? -ant::IS2 : ant::IS2;
  c.ayl = ant::IS2;
  if (k == 0 || k == 3) { c.lo2 = ant::ANK_LO; c.hi2 = ant::ANK_HI; }
  else { c.lo2 = -ant::ANK_HI; c.hi2 = -ant::ANK_LO; }
  return c;
}

// leg geometry in world axes, relative to the torso origin
struct LegKin {
  V3 ex, ey, ez;  // torso frame axes
  V3 rh;          // hip point - O
  V3 r1;          // aux COM - hip   (ankle - hip = 2 r1)
  V3 a2;          // ankle axis (hip axis a1 = ez)
  V3 r2;          // foot COM - ankle (tip - ankle = 2 r2)
  V3 zf;          // foot frame z axis
};

__device__ __forceinline__ LegKin leg_fk(const AntLane& s, const LegConst& c) {
  LegKin K;
  float x = s.qx, y = s.qy, z = s.qz, w = s.qw;
  K.ex = mk(1.f - 2.f * (y * y + z * z), 2.f * (x * y + z * w), 2.f * (x * z - y * w));
  K.ey = mk(2.f * (x * y - z * w), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + x * w));
  K.ez = mk(2.f * (x * z + y * w), 2.f * (y * z - x * w), 1.f - 2.f * (x * x + y * y));
  V3 d0 = c.sx * K.ex + c.sy * K.ey;
  K.rh = 0.2f * d0;
  float s1, c1, s2, c2;
  __sincosf(s.q1, &s1, &c1);  // |q| < 2 rad: MUFU sin/cos, abs error ~5e-7
  __sincosf(s.q2, &s2, &c2);
  V3 auxx = c1 * K.ex + s1 * K.ey, auxy = c1 * K.ey - s1 * K.ex;
  V3 d1 = c.sx * auxx + c.sy * auxy;
  K.r1 = 0.1f * d1;
  K.a2 = c.axl * auxx + c.ayl * auxy;
  float wsgn = c.axl * c.sy - c.ayl * c.sx;      // a2 x d1 = wsgn * ez
  V3 b = c.ayl * auxx - c.axl * auxy;            // a2 x ez
  V3 d2 = c2 * d1 + (s2 * wsgn) * K.ez;          // Rodrigues, a2 _|_ d1
  K.zf = c2 * K.ez + s2 * b;                     // a2 _|_ ez
  K.r2 = 0.2f * d2;
  return K;
}

// 6x6 symmetric index (i <= j)
__host__ __device__ constexpr int sym6(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }

struct LegDyn {
  float mi11, mi12, mi22;  // Mll^-1
  float il11, l21, il22;   // Mll = Ll Ll^T: 1/l11, l21, 1/l22
  float G1[6], G2[6];      // base<-joint coupling columns [torque about O; force]
  float K0[6], K1[6];      // (Mll^-1 G^T) rows = G Mll^-1 columns
  float Li[21];            // L^-1 (lower, packed row-major: Li[i*(i+1)/2 + j], j <= i) of S = L L^T
};

// x = S^-1 b  via  L^-T (L^-1 b)
__device__ __forceinline__ void sinv_mul(const float* __restrict__ Li, const float b[6], float x[6]) {
  float t[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j <= i; j++) a = fmaf(Li[i * (i + 1) / 2 + j], b[j], a);
    t[i] = a;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float a = 0.f;
#pragma unroll
    for (int j = i; j < 6; j++) a = fmaf(Li[j * (j + 1) / 2 + i], t[j], a);
    x[i] = a;
  }
}

__device__ __forceinline__ float clampf(float x, float lim) { return fminf(fmaxf(x, -lim), lim); }

// Bullet btPlaneSpace1
__device__ __forceinline__ void plane_space(V3 n, V3& p, V3& q) {
  if (fabsf(n.z) > 0.7071067811865475244f) {
    float a = n.y * n.y + n.z * n.z, k = rsqrt_ftz(a);
    p = mk(0.f, -n.z * k, n.y * k);
    q = mk(a * k, -n.x * p.z, n.x * p.y);
  } else {
    float a = n.x * n.x + n.y * n.y, k = rsqrt_ftz(a);
    p = mk(-n.y * k, n.x * k, 0.f);
    q = mk(-n.z * p.y, n.z * p.x, a * k);
  }
}

struct SubstepParams {
  float h, inv_h, g, kl, ka, erp_c, erp_l, mu, max_imp, vmax, margin, gz;
  float wx, wy;  // wall inner faces at +-wx, +-wy
  int has_walls, has_box, iters;
  float blo[3], bhi[3];
  int item_contacts, n_items;  // cube colliders (AntGather only)
  float mu_item, item_half, item_z;
};

__device__ __forceinline__ SubstepParams make_params(const hrl_config& cfg) {
  SubstepParams p;
  p.h = cfg.dt / (float)cfg.substeps; p.inv_h = 1.0f / p.h; p.g = cfg.gravity; p.kl = cfg.lin_damping; p.ka = cfg.ang_damping;
  p.erp_c = cfg.contact_erp; p.erp_l = cfg.limit_erp; p.mu = cfg.friction; p.max_imp = cfg.limit_max_impulse;
  p.vmax = cfg.max_coord_vel; p.margin = cfg.contact_margin; p.gz = cfg.ground_z;
  p.wx = cfg.world_size[0] * 0.5f - 0.05f; p.wy = cfg.world_size[1] * 0.5f - 0.05f;
  p.has_walls = cfg.has_walls; p.has_box = cfg.has_box; p.iters = cfg.solver_iters;
  p.item_contacts = cfg.item_contacts && cfg.env_kind == HRL_ANT_GATHER; p.n_items = cfg.n_food + cfg.n_poison;
  p.mu_item = cfg.item_friction; p.item_half = cfg.item_half; p.item_z = cfg.item_z;
#pragma unroll
  for (int i = 0; i < 3; i++) { p.blo[i] = cfg.box_lo[i]; p.bhi[i] = cfg.box_hi[i]; }
  return p;
}

#define CAND(c, f) cands[((c) * HRL_CAND_F + (f)) * 32 + lane]

// exp(w h) of the quaternion update for |w| h > pi/4 (needs max_coord_vel > 110): Bullet's literal path with its angle
// cap, (sin(y)/|w|, cos(y)).  Never taken at the default clamp; out of line (results by value, in registers) so that
// sinf / cosf stay off the sub-step loop's instruction footprint.
__device__ __noinline__ float2 quat_exp_literal(float w2, float h) {
  float ang = sqrtf(w2);
  if (ang * h > 0.25f * 3.14159265358979323846f) ang = 0.25f * 3.14159265358979323846f / h;
  return make_float2(sinf(0.5f * ang * h) / ang, cosf(0.5f * ang * h));
}
// Append one contact candidate of this lane (contact point on the robot relative to O, normal, distance, body).
__device__ __forceinline__ int add_cand(float* __restrict__ cands, int lane, int nC, V3 crel, float r, V3 n, float dist, float body) {
  HRL_CHECK(nC >= 0 && nC <= HRL_MAXC && lane >= 0 && lane < 32);
  if (nC < HRL_MAXC) {
    const V3 Prel = crel - r * n;
    CAND(nC, 0) = Prel.x; CAND(nC, 1) = Prel.y; CAND(nC, 2) = Prel.z;
    CAND(nC, 3) = n.x; CAND(nC, 4) = n.y; CAND(nC, 5) = n.z;
    CAND(nC, 6) = dist; CAND(nC, 7) = body;
    nC++;
  }
  return nC;
}
// Sphere (centre c = O + crel, radius r) against the four wall planes at +-wx, +-wy, order +x -x +y -y like the oracle.
__device__ __noinline__ int sphere_vs_walls(V3 c, V3 crel, float r, float body, float wx, float wy, float margin,
                                            float* __restrict__ cands, int lane, int nC) {
  float d;
  d = wx - c.x - r; if (d < margin) nC = add_cand(cands, lane, nC, crel, r, mk(-1.f, 0.f, 0.f), d, body);
  d = c.x + wx - r; if (d < margin) nC = add_cand(cands, lane, nC, crel, r, mk(1.f, 0.f, 0.f), d, body);
  d = wy - c.y - r; if (d < margin) nC = add_cand(cands, lane, nC, crel, r, mk(0.f, -1.f, 0.f), d, body);
  d = c.y + wy - r; if (d < margin) nC = add_cand(cands, lane, nC, crel, r, mk(0.f, 1.f, 0.f), d, body);
  return nC;
}
// Sphere against an axis-aligned box (maze box, food / poison cubes): closest point outside, nearest face inside.
// Returns the new candidate count, with bit 8 set when the sphere is within the margin (a contact POINT exists even if
// the lane's candidate list is full - what the contact-based pickup counts).
__device__ __forceinline__ int sphere_vs_aabb_inl(V3 c, V3 crel, float r, float body, float lox, float loy, float loz, float hix,
                                                  float hiy, float hiz, float margin, float* __restrict__ cands, int lane, int nC) {
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
  V3 n; float dist;
  if (!inside) {
    const V3 d = mk(cc[0] - qq[0], cc[1] - qq[1], cc[2] - qq[2]);
    const float len = sqrtf(dot(d, d));
    n = (1.0f / len) * d; dist = len - r;
  } else {
    float best = 1e30f; int bi = 0; float bs = 1.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float dl = cc[i] - lo[i], dh = hi[i] - cc[i];
      if (dl < best) { best = dl; bi = i; bs = -1.f; }
      if (dh < best) { best = dh; bi = i; bs = 1.f; }
    }
    n = mk(bi == 0 ? bs : 0.f, bi == 1 ? bs : 0.f, bi == 2 ? bs : 0.f);
    dist = -best - r;
  }
  if (dist < margin) nC = add_cand(cands, lane, nC, crel, r, n, dist, body) | 0x100;
  return nC;
}
__device__ __noinline__ int sphere_vs_aabb(V3 c, V3 crel, float r, float body, float lox, float loy, float loz, float hix,
                                           float hiy, float hiz, float margin, float* __restrict__ cands, int lane, int nC) {
  return sphere_vs_aabb_inl(c, crel, r, body, lox, loy, loz, hix, hiy, hiz, margin, cands, lane, nC);
}

// Maze box: the cylinder part of one leg's three capsules against the box's four vertical edges (the end-spheres only
// cover contacts at a capsule end; for a segment outside a convex rectangle the other closest pair is (rectangle
// corner, segment interior); planar because the legs never reach the box's top).  Same order as the oracle
// (detect_contacts): foot, aux, leg capsule x corners (lo,lo) (hi,lo) (lo,hi) (hi,hi).  Appends to the lane's
// candidate list, returns the new count.
__device__ __noinline__ int capsules_vs_box_edges(V3 O, V3 rh, V3 r_ank, V3 r_tip, float lox, float loy, float loz, float hix,
                                                  float hiy, float hiz, float margin, float* __restrict__ cands, int lane, int nC) {
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
      if (ex * sgx < 0.f || ey * sgy < 0.f) continue;  // not in the corner's Voronoi region: a face is closer
      const float e2 = ex * ex + ey * ey;
      if (!(e2 > 0.f)) continue;
      const float el = sqrtf(e2), dist = el - ant::R_CAPS;
      if (dist < margin && nC < HRL_MAXC) {
        const float nx = ex / el, ny = ey / el;
        CAND(nC, 0) = Q.x - ant::R_CAPS * nx; CAND(nC, 1) = Q.y - ant::R_CAPS * ny; CAND(nC, 2) = Q.z;
        CAND(nC, 3) = nx; CAND(nC, 4) = ny; CAND(nC, 5) = 0.f;
        CAND(nC, 6) = dist; CAND(nC, 7) = (float)(2 - cap);
        nC++;
      }
    }
  }
  return nC;
}

// Food / poison cubes: the CYLINDER part of one leg's three capsules against the candidate cubes (the end-spheres only
// cover contacts at a capsule end; a leg lying across a cube's edge touches it in between).  Same algorithm and order as
// the oracle (capsule_interior_vs_box, detect_contacts): cubes in index order, capsules foot, aux, leg; the closest
// point of the segment's interior is bracketed by 24 bisection steps on the (monotone) derivative of the squared
// point-box distance; normal and distance there are those of a sphere of the capsule radius.  Candidate (slot, field)
// lives at cbase[(slot * HRL_CAND_F + field) * stride]; slots slot0, slot0 + 1, ... are filled while < HRL_MAXC.
// Returns the number of contact points found (stored or not).  Rare: out of line.
__device__ __forceinline__ float seg_box_grad(const float a[3], const float d[3], float h, float t) {
  float g = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float x = fmaf(t, d[i], a[i]);
    const float e = x > h ? x - h : (x < -h ? x + h : 0.f);
    g = fmaf(e, d[i], g);
  }
  return g;
}
__device__ __noinline__ int capsules_vs_cubes(V3 O, V3 rh, V3 r_ank, V3 r_tip, unsigned imask, const float* __restrict__ ixy, float half,
                                              float iz, float margin, float* __restrict__ cbase, int stride, int slot0,
                                              unsigned long long* itouch, bool count_touch) {
  int n = 0;
  const float rr = half + ant::R_CAPS + margin;
  for (unsigned m = imask; m; m &= m - 1) {
    const int gi = __ffs(m) - 1;
    const float bx = ixy[2 * gi] - O.x, by = ixy[2 * gi + 1] - O.y, bz = iz - O.z;  // cube centre relative to the torso origin
#pragma unroll 1
    for (int cap = 0; cap < 3; cap++) {
      const V3 A = cap == 0 ? r_ank : (cap == 1 ? rh : mk(0.f, 0.f, 0.f));
      const V3 B = cap == 0 ? r_tip : (cap == 1 ? r_ank : rh);
      if (bx < fminf(A.x, B.x) - rr || bx > fmaxf(A.x, B.x) + rr || by < fminf(A.y, B.y) - rr || by > fmaxf(A.y, B.y) + rr) continue;
      const float a[3] = {A.x - bx, A.y - by, A.z - bz}, d[3] = {B.x - A.x, B.y - A.y, B.z - A.z};   // relative to the cube centre
      if (!(seg_box_grad(a, d, half, 0.f) < 0.f && seg_box_grad(a, d, half, 1.f) > 0.f)) continue;  // minimum at an end: the spheres' business
      float lo = 0.f, hi = 1.f;
#pragma unroll 1
      for (int it = 0; it < 24; it++) {
        const float mid = 0.5f * (lo + hi);
        if (seg_box_grad(a, d, half, mid) < 0.f) lo = mid; else hi = mid;
      }
      const float t = 0.5f * (lo + hi);
      const float q[3] = {fmaf(t, d[0], a[0]), fmaf(t, d[1], a[1]), fmaf(t, d[2], a[2])};
      float e[3];
      bool inside = true;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        e[i] = q[i] > half ? q[i] - half : (q[i] < -half ? q[i] + half : 0.f);
        inside = inside && e[i] == 0.f;
      }
      V3 nrm; float dist;
      if (!inside) {
        const float len = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        nrm = mk(e[0] / len, e[1] / len, e[2] / len); dist = len - ant::R_CAPS;
      } else {  // the segment enters the cube: exit through the nearest face (as sphere_vs_aabb_inl)
        float best = 1e30f; int bi = 0; float bs = 1.f;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const float dl = q[i] + half, dh = half - q[i];
          if (dl < best) { best = dl; bi = i; bs = -1.f; }
          if (dh < best) { best = dh; bi = i; bs = 1.f; }
        }
        nrm = mk(bi == 0 ? bs : 0.f, bi == 1 ? bs : 0.f, bi == 2 ? bs : 0.f); dist = -best - ant::R_CAPS;
      }
      if (!(dist < margin)) continue;
      if (count_touch) HRL_TOUCH_ADD(itouch, gi);
      const int slot = slot0 + n;
      n++;
      if (slot < HRL_MAXC) {
        float* c = cbase + slot * HRL_CAND_F * stride;
        c[0] = q[0] + bx - ant::R_CAPS * nrm.x; c[stride] = q[1] + by - ant::R_CAPS * nrm.y; c[2 * stride] = q[2] + bz - ant::R_CAPS * nrm.z;
        c[3 * stride] = nrm.x; c[4 * stride] = nrm.y; c[5 * stride] = nrm.z;
        c[6 * stride] = dist; c[7 * stride] = (float)(2 - cap) + 4.f;   // link class (foot 2, aux 1, leg 0) + 4: cube friction
      }
    }
  }
  return n;
}

// Whiten one constraint row of leg k and store it at visit position `pos` of this env's row buffer.
//   Jb~ = JB - K [j1 j2]^T (leg eliminated), z = L^-1 Jb~, y = Ll^-1 [j1 j2]^T, diag = |z|^2 + |y|^2.
template <bool COMPACT = false>
__device__ __forceinline__ void emit_row(float4* __restrict__ rb, int pos, int k,
                                         const LegDyn& D, const float JB[6], float j1, float j2, const float ub[6],
                                         float u1, float u2, float pen, float erp, float inv_h, bool positional,
                                         bool fast = false, int slot = 0) {
  HRL_CHECK(fast || (pos >= 0 && pos < HRL_ROWS_ENV && pos != HRL_ROW_ZERO && pos != HRL_ROW_ZERO + 1));
  HRL_CHECK(!fast || (slot >= 0 && slot < HRL_DS_P));
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
  if (!COMPACT && HRL_DELASSUS && fast) {  // transposed, into static position `slot` (cleared beforehand: the other legs' components stay 0)
    float* f = reinterpret_cast<float*>(rb);
#pragma unroll
    for (int i = 0; i < 6; i++) f[HRL_DS_AT + i * HRL_DS_P + slot] = z[i];
    f[HRL_DS_AT + (6 + 2 * k) * HRL_DS_P + slot] = y0;
    f[HRL_DS_AT + (7 + 2 * k) * HRL_DS_P + slot] = y1;
    f[HRL_DS_DINV + slot] = dinv;
    f[HRL_DS_RHS + slot] = (posErr + velErr) * dinv;
    reinterpret_cast<int*>(f)[HRL_DS_LEG + slot] = k;
    return;
  }
  if (COMPACT) {
    float4* r = rb + pos * 3;
    r[0] = make_float4(z[0], z[1], z[2], z[3]);
    r[1] = make_float4(z[4], z[5], y0, y1);
    r[2] = make_float4(dinv, (posErr + velErr) * dinv, __int_as_float(k), 0.f);
    return;
  }
  float4* r = rb + pos * 4;
  r[0] = make_float4(z[0], z[1], z[2], z[3]);
  r[1] = make_float4(z[4], z[5], k == 0 ? y0 : 0.f, k == 0 ? y1 : 0.f);
  r[2] = make_float4(k == 1 ? y0 : 0.f, k == 1 ? y1 : 0.f, k == 2 ? y0 : 0.f, k == 2 ? y1 : 0.f);
  r[3] = make_float4(k == 3 ? y0 : 0.f, k == 3 ? y1 : 0.f, dinv, (posErr + velErr) * dinv);
}

// ---- Blackwell packed fp32 (FFMA2 / FMUL2 / FADD2): two IEEE-rn fp32 operations per instruction ----
#define HRL_U64(v) reinterpret_cast<unsigned long long&>(v)
#ifndef HRL_SCALAR_FMA
#define HRL_SCALAR_FMA 0   // 1: plain fp32 instructions instead of the packed ones (A/B only)
#endif
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  if (HRL_SCALAR_FMA) return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
  float2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(HRL_U64(r)) : "l"(HRL_U64(a)), "l"(HRL_U64(b)), "l"(HRL_U64(c)));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  if (HRL_SCALAR_FMA) return make_float2(a.x * b.x, a.y * b.y);
  float2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(HRL_U64(r)) : "l"(HRL_U64(a)), "l"(HRL_U64(b)));
  return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  if (HRL_SCALAR_FMA) return make_float2(a.x + b.x, a.y + b.y);
  float2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(HRL_U64(r)) : "l"(HRL_U64(a)), "l"(HRL_U64(b)));
  return r;
}

struct Row { float2 p[7]; float dinv, rhs; };  // a[0..13] as 7 pairs, 1/diag, rhs/diag
template <bool COMPACT = false>
__device__ __forceinline__ Row ld_row(const float4* __restrict__ rb, int r) {
  Row R;
  HRL_CHECK(r >= 0 && r < HRL_ROWS_ENV);
  if (COMPACT) {
    const float4 q0 = rb[r * 3], q1 = rb[r * 3 + 1], q2 = rb[r * 3 + 2];
    const int leg = __float_as_int(q2.z);
    const float2 y = make_float2(q1.z, q1.w), o = make_float2(0.f, 0.f);
    R.p[0] = make_float2(q0.x, q0.y); R.p[1] = make_float2(q0.z, q0.w); R.p[2] = make_float2(q1.x, q1.y);
    R.p[3] = leg == 0 ? y : o; R.p[4] = leg == 1 ? y : o; R.p[5] = leg == 2 ? y : o; R.p[6] = leg == 3 ? y : o;
    R.dinv = q2.x; R.rhs = q2.y;
    return R;
  }
  const float4 q0 = rb[r * 4], q1 = rb[r * 4 + 1], q2 = rb[r * 4 + 2], q3 = rb[r * 4 + 3];
  R.p[0] = make_float2(q0.x, q0.y); R.p[1] = make_float2(q0.z, q0.w); R.p[2] = make_float2(q1.x, q1.y);
  R.p[3] = make_float2(q1.z, q1.w); R.p[4] = make_float2(q2.x, q2.y); R.p[5] = make_float2(q2.z, q2.w);
  R.p[6] = make_float2(q3.x, q3.y); R.dinv = q3.z; R.rhs = q3.w;
  return R;
}
__device__ __forceinline__ float dot14(const float2 a[7], const float2 b[7]) {
  // four packed accumulators: dependent depth mul, fma, add2, add2, add
  float2 s0 = mul2(a[0], b[0]), s1 = mul2(a[1], b[1]), s2 = mul2(a[2], b[2]), s3 = mul2(a[3], b[3]);
  s0 = fma2(a[4], b[4], s0); s1 = fma2(a[5], b[5], s1); s2 = fma2(a[6], b[6], s2);
  const float2 t = add2(add2(s0, s1), add2(s2, s3));
  return t.x + t.y;
}
__device__ __forceinline__ void axpy14(float2 y[7], const float2 a[7], float s) {
  const float2 ss = make_float2(s, s);
#pragma unroll
  for (int i = 0; i < 7; i++) y[i] = fma2(a[i], ss, y[i]);
}

// One visit of a single (non-friction) row: clamp the impulse to [0, hi], apply the delta.
// Idle visits use an all-zero row (HRL_ROW_ZERO; a = 0, dinv = rhs = 0, impulse 0): dl = 0 falls out
// of the arithmetic, no predicate needed.
template <bool HAS_HI>
__device__ __forceinline__ float single_visit(float* __restrict__ lp, float2 dv[7], const Row& R, float lam, float hi) {
  // new impulse = clamp(lam + rhs' - (a . dv') / (a . a)): lam + rhs' is formed off the dependency chain, which is
  // dot -> fma -> max (-> min) -> sub -> axpy
  float nl = fmaxf(fmaf(-dot14(R.p, dv), R.dinv, lam + R.rhs), 0.f);
  if (HAS_HI) nl = fminf(nl, hi);
  *lp = nl;
  axpy14(dv, R.p, nl - lam);
  return nl - lam;
}
// The same visit with the row velocity a . dv' handed in (look-ahead, see the sweep): returns nothing, applies the delta.
template <bool HAS_HI>
__device__ __forceinline__ void single_visit_pre(float* __restrict__ lp, float2 dv[7], const Row& R, float lam, float hi, float d) {
  float nl = fmaxf(fmaf(-d, R.dinv, lam + R.rhs), 0.f);
  if (HAS_HI) nl = fminf(nl, hi);
  *lp = nl;
  axpy14(dv, R.p, nl - lam);
}
// One friction pair with the implicit cone |f| <= mu * lambda_n; skipped (impulses kept) when the
// normal impulse is zero, like Bullet.  c = (lambda_t1, lambda_t2, lambda_n, mu) of the contact slot at *cp.
__device__ __forceinline__ void pair_visit(float4* __restrict__ cp, float2 dv[7], const Row& A, const Row& B, const float4 c) {
  float na = fmaf(-dot14(A.p, dv), A.dinv, c.x + A.rhs);
  float nb = fmaf(-dot14(B.p, dv), B.dinv, c.y + B.rhs);
  const float lim = c.w * c.z, len2 = fmaf(na, na, nb * nb);
  const float sc = (len2 > lim * lim) ? lim * rsqrt_ftz(fmaxf(len2, 1e-30f)) : 1.f;
  const bool on = c.z > 0.f;
  na = on ? na * sc : c.x; nb = on ? nb * sc : c.y;
  *reinterpret_cast<float2*>(cp) = make_float2(na, nb);
  axpy14(dv, A.p, na - c.x); axpy14(dv, B.p, nb - c.y);
}

// ---- Delassus-space sweep: helpers (positions and layout at the top of this file) ----
// Sweep state of one lane: targets / impulses of the 8 limit rows (the same in the 4 lanes of the env) and of the
// lane's own contact (x normal, y friction 1, z friction 2, w pad).
struct DsState {
  float2 tL[4], lL[4];
  float4 tC, lC;
  float mu;
};
template <int S>
__device__ __forceinline__ float& ds_at(float2 a[4]) { return (S & 1) ? a[S >> 1].y : a[S >> 1].x; }
// t_j -= What[r][j] * dl: the 8 limit targets (2 broadcast LDS.128) and the lane's own contact (1 LDS.128)
__device__ __forceinline__ void ds_apply(DsState& st, const float* __restrict__ W, int r, int k, float dl) {
  const float4* __restrict__ w = reinterpret_cast<const float4*>(W + r * HRL_DS_P);
  const float2 d = make_float2(-dl, -dl);
  const float4 q0 = w[0], q1 = w[1], qc = w[2 + k];
  st.tL[0] = fma2(make_float2(q0.x, q0.y), d, st.tL[0]);
  st.tL[1] = fma2(make_float2(q0.z, q0.w), d, st.tL[1]);
  st.tL[2] = fma2(make_float2(q1.x, q1.y), d, st.tL[2]);
  st.tL[3] = fma2(make_float2(q1.z, q1.w), d, st.tL[3]);
  float2 a = make_float2(st.tC.x, st.tC.y), b = make_float2(st.tC.z, st.tC.w);
  a = fma2(make_float2(qc.x, qc.y), d, a);
  b = fma2(make_float2(qc.z, qc.w), d, b);
  st.tC = make_float4(a.x, a.y, b.x, b.y);
}
// joint-limit row S: every lane holds its target, no communication
template <int S>
__device__ __forceinline__ void ds_limit(DsState& st, const float* __restrict__ W, int k, float hi) {
  float& lam = ds_at<S>(st.lL);
  const float nl = fminf(fmaxf(ds_at<S>(st.tL), 0.f), hi);
  const float dl = nl - lam;
  lam = nl;
  ds_apply(st, W, S, k, dl);
}
// normal of contact C: lane C computes, one shuffle tells the others
template <int C>
__device__ __forceinline__ void ds_normal(DsState& st, const float* __restrict__ W, int k) {
  const float nl = fmaxf(st.tC.x, 0.f);
  const float dl = __shfl_sync(HRL_FULL_MASK, nl - st.lC.x, C, 4);
  if (k == C) st.lC.x = nl;
  ds_apply(st, W, 8 + 3 * C, k, dl);
}
// friction pair of contact C with the implicit cone |f| <= mu * lambda_n; kept as it is while lambda_n = 0 (pair_visit)
template <int C>
__device__ __forceinline__ void ds_pair(DsState& st, const float* __restrict__ W, int k) {
  float na = st.tC.y, nb = st.tC.z;
  const float lim = st.mu * st.lC.x, len2 = fmaf(na, na, nb * nb);
  const float sc = (len2 > lim * lim) ? lim * rsqrt_ftz(fmaxf(len2, 1e-30f)) : 1.f;
  const bool on = st.lC.x > 0.f;
  na = on ? na * sc : st.lC.y; nb = on ? nb * sc : st.lC.z;
  const float dla = __shfl_sync(HRL_FULL_MASK, na - st.lC.y, C, 4), dlb = __shfl_sync(HRL_FULL_MASK, nb - st.lC.z, C, 4);
  if (k == C) { st.lC.y = na; st.lC.z = nb; }
  ds_apply(st, W, 9 + 3 * C, k, dla);
  ds_apply(st, W, 10 + 3 * C, k, dlb);
}
// the 3 rows of one contact (same leg: lrc = that leg's two components of AT) against the 4 columns of group g:
// acc[i] = sum_m a_i[m] * AT[m][4g .. 4g+3] - 8 broadcast LDS.128 feed 48 FFMA2 in 6 independent chains
__device__ __forceinline__ void ds_gram_group3(const float* __restrict__ f, const float2 (*zz)[8], const float* __restrict__ lrc, int g,
                                               float2 (*acc)[2]) {
#pragma unroll
  for (int i = 0; i < 3; i++) { acc[i][0] = make_float2(0.f, 0.f); acc[i][1] = make_float2(0.f, 0.f); }
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const float4 w = *reinterpret_cast<const float4*>((m < 6 ? f + HRL_DS_AT + m * HRL_DS_P : lrc + (m - 6) * HRL_DS_P) + 4 * g);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      acc[i][0] = fma2(zz[i][m], make_float2(w.x, w.y), acc[i][0]);
      acc[i][1] = fma2(zz[i][m], make_float2(w.z, w.w), acc[i][1]);
    }
  }
}
// What from the transposed rows.  Lane k computes the 3 rows of contact k against all columns (and, transposed, the
// contact's columns of the limit rows: What[j][c] = W[c][j] / diag_c), then limit rows k and k + 4 against the limit columns.
// gmask: warp-uniform mask of the column groups that hold a row in ANY env of the warp (0: limits 0-3, 1: limits 4-7, 2 + c: contact c).
__device__ __forceinline__ void ds_gram(float* __restrict__ f, int k, unsigned gmask) {
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gmask & ~3u) {  // ---- contact rows
    const int p0 = 8 + 4 * k;
    const int leg = reinterpret_cast<const int*>(f)[HRL_DS_LEG + p0];
    const float* lrc = f + HRL_DS_AT + (6 + 2 * leg) * HRL_DS_P;
    float2 zz[3][8];
#pragma unroll
    for (int m = 0; m < 8; m++) {
      const float4 v = *reinterpret_cast<const float4*>((m < 6 ? f + HRL_DS_AT + m * HRL_DS_P : lrc + (m - 6) * HRL_DS_P) + p0);
      zz[0][m] = make_float2(v.x, v.x); zz[1][m] = make_float2(v.y, v.y); zz[2][m] = make_float2(v.z, v.z);
    }
    const float4 dme = *reinterpret_cast<const float4*>(f + HRL_DS_DINV + p0);  // 1 / diag of this contact's rows
    float4* __restrict__ wout = reinterpret_cast<float4*>(f + HRL_DS_W + (8 + 3 * k) * HRL_DS_P);
#pragma unroll 1
    for (int g = 0; g < HRL_DS_P / 4; g++) {
      if (!((gmask >> g) & 1u)) {
#pragma unroll
        for (int i = 0; i < 3; i++) wout[i * (HRL_DS_P / 4) + g] = zero4;
        if (g < 2) {
#pragma unroll
          for (int j = 0; j < 4; j++) *reinterpret_cast<float4*>(f + HRL_DS_W + (4 * g + j) * HRL_DS_P + p0) = zero4;
        }
        continue;
      }
      float2 acc[3][2];
      ds_gram_group3(f, zz, lrc, g, acc);
      if (g < 2) {  // transposed: the limit rows' columns of this contact
        const float wj[3][4] = {{acc[0][0].x, acc[0][0].y, acc[0][1].x, acc[0][1].y}, {acc[1][0].x, acc[1][0].y, acc[1][1].x, acc[1][1].y},
                                {acc[2][0].x, acc[2][0].y, acc[2][1].x, acc[2][1].y}};
#pragma unroll
        for (int j = 0; j < 4; j++)
          *reinterpret_cast<float4*>(f + HRL_DS_W + (4 * g + j) * HRL_DS_P + p0) = make_float4(wj[0][j] * dme.x, wj[1][j] * dme.y, wj[2][j] * dme.z, 0.f);
      }
      const float4 d = *reinterpret_cast<const float4*>(f + HRL_DS_DINV + 4 * g);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        acc[i][0] = mul2(acc[i][0], make_float2(d.x, d.y)); acc[i][1] = mul2(acc[i][1], make_float2(d.z, d.w));
      }
      if (g == 2 + k) { acc[0][0].x = 0.f; acc[1][0].y = 0.f; acc[2][1].x = 0.f; }  // zero diagonal
#pragma unroll
      for (int i = 0; i < 3; i++) wout[i * (HRL_DS_P / 4) + g] = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
    }
  } else if (gmask & 3u) {  // no contact anywhere in the warp: the limit rows' contact columns are never used, but must be finite
#pragma unroll
    for (int j = 0; j < 8; j++) *reinterpret_cast<float4*>(f + HRL_DS_W + j * HRL_DS_P + 8 + 4 * k) = zero4;
  }
  if (gmask & 3u) {  // ---- limit rows k and k + 4 against the limit columns
    const int leg0 = reinterpret_cast<const int*>(f)[HRL_DS_LEG + k], leg1 = reinterpret_cast<const int*>(f)[HRL_DS_LEG + 4 + k];
    const float* const lr[2] = {f + HRL_DS_AT + (6 + 2 * leg0) * HRL_DS_P, f + HRL_DS_AT + (6 + 2 * leg1) * HRL_DS_P};
    float2 zz[2][8];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int m = 0; m < 8; m++) {
        const float v = (m < 6 ? f + HRL_DS_AT + m * HRL_DS_P : lr[i] + (m - 6) * HRL_DS_P)[4 * i + k];
        zz[i][m] = make_float2(v, v);
      }
#pragma unroll 1
    for (int g = 0; g < 2; g++) {
      float4* __restrict__ w0 = reinterpret_cast<float4*>(f + HRL_DS_W + k * HRL_DS_P) + g;
      float4* __restrict__ w1 = reinterpret_cast<float4*>(f + HRL_DS_W + (4 + k) * HRL_DS_P) + g;
      if (!((gmask >> g) & 1u)) { *w0 = zero4; *w1 = zero4; continue; }
      float2 acc[2][2];
      // (two different legs: the leg part of each row reads its own two components)
#pragma unroll
      for (int i = 0; i < 2; i++) { acc[i][0] = make_float2(0.f, 0.f); acc[i][1] = make_float2(0.f, 0.f); }
#pragma unroll
      for (int m = 0; m < 8; m++) {
        if (m < 6) {  // base components: one load feeds both rows; leg components: per row
          const float4 w = *reinterpret_cast<const float4*>(f + HRL_DS_AT + m * HRL_DS_P + 4 * g);
#pragma unroll
          for (int i = 0; i < 2; i++) {
            acc[i][0] = fma2(zz[i][m], make_float2(w.x, w.y), acc[i][0]);
            acc[i][1] = fma2(zz[i][m], make_float2(w.z, w.w), acc[i][1]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const float4 w = *reinterpret_cast<const float4*>(lr[i] + (m - 6) * HRL_DS_P + 4 * g);
            acc[i][0] = fma2(zz[i][m], make_float2(w.x, w.y), acc[i][0]);
            acc[i][1] = fma2(zz[i][m], make_float2(w.z, w.w), acc[i][1]);
          }
        }
      }
      const float4 d = *reinterpret_cast<const float4*>(f + HRL_DS_DINV + 4 * g);
#pragma unroll
      for (int i = 0; i < 2; i++) {
        acc[i][0] = mul2(acc[i][0], make_float2(d.x, d.y)); acc[i][1] = mul2(acc[i][1], make_float2(d.z, d.w));
        if (g == i) {  // zero diagonal: row 4 i + k, column k of group i
          if (k == 0) acc[i][0].x = 0.f;
          if (k == 1) acc[i][0].y = 0.f;
          if (k == 2) acc[i][1].x = 0.f;
          if (k == 3) acc[i][1].y = 0.f;
        }
      }
      *w0 = make_float4(acc[0][0].x, acc[0][0].y, acc[0][1].x, acc[0][1].y);
      *w1 = make_float4(acc[1][0].x, acc[1][0].y, acc[1][1].x, acc[1][1].y);
    }
  }
}

// One internal step of h = dt/substeps.  `rows`/`cands` point at this WARP's shared memory
// (rows: HRL_ROWS_FLOATS_PER_WARP row floats followed by HRL_LAM_FLOATS_PER_WARP impulses).
// feet_ground: bit set when this leg's foot link (tip or ankle sphere) has a floor manifold at
// the START of the sub-step (collision detection precedes the dynamics, like Bullet).
// ITEMS: compile the food / poison cube colliders in (AntGather).  it_x / it_y: this lane's 4 items (4k..4k+3);
// iscr: HRL_ITEM_SCRATCH_FLOATS floats of this warp; count_touch: tally contact points per cube (last sub-step).
template <bool ITEMS, bool COMPACT = false>
__device__ __forceinline__ void ant_substep(AntLane& s, const SubstepParams& P, const LegConst& lc, float tau1,
                                            float tau2, float* __restrict__ rows, float* __restrict__ cands,
                                            int lane, int k, int es, int& feet_ground, int& stat_contacts, int& stat_limits,
                                            const float* it_x = nullptr, const float* it_y = nullptr, float* iscr = nullptr,
                                            bool count_touch = false, int* trips = nullptr, float* dbg = nullptr) {
  const LegKin K = leg_fk(s, lc);
  const V3 a1 = K.ez, a2 = K.a2;
  const V3 r_ac = K.rh + K.r1;           // aux COM - O
  const V3 r_ank = K.rh + 2.f * K.r1;    // ankle - O
  const V3 rf1 = 2.f * K.r1 + K.r2;      // foot COM - hip
  const V3 r_fc = K.rh + rf1;            // foot COM - O
  const V3 r_tip = r_ank + 2.f * K.r2;   // foot tip - O

  // ---------------- contacts: spheres vs ground / walls / maze box (start-of-step poses) ----------
  int nC = 0;
  feet_ground = 0;
  {
    // candidate order inside the group: [torso sphere (lane 0)], tip (foot), ankle (aux), hip (torso);
    // per sphere: ground, walls +x -x +y -y, box.  Rolled on purpose (code size); the four wall tests
    // hide behind one comparison because a wall contact is rare.
    // cube colliders: cubes within reach of the robot (torso-distance cull: longest reach 1.131 m + half + r + margin
    // per axis) are staged in shared memory with a 16-bit candidate mask per env; usually the mask is empty
    unsigned imask = 0;
    float* ixy = nullptr;
    unsigned long long* itouch = nullptr;
    if (ITEMS && P.item_contacts) {
      ixy = iscr + es * 32; itouch = reinterpret_cast<unsigned long long*>(iscr + HRL_EPW * 32) + 2 * es;
      const float reach = 1.1314f + 1.4143f * (P.item_half + ant::R_CAPS + P.margin);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int gi = 4 * k + i;
        HRL_CHECK(es >= 0 && es < HRL_EPW && gi < 16);
        const float dx = it_x[i] - s.O.x, dy = it_y[i] - s.O.y;
        if (gi < P.n_items && dx * dx + dy * dy < reach * reach) imask |= 1u << gi;
        ixy[2 * gi] = it_x[i]; ixy[2 * gi + 1] = it_y[i];
      }
      if (count_touch && k == 0) { itouch[0] = 0ull; itouch[1] = 0ull; }
      imask |= __shfl_xor_sync(HRL_FULL_MASK, imask, 1);
      imask |= __shfl_xor_sync(HRL_FULL_MASK, imask, 2);
      __syncwarp();
    }
#pragma unroll 1
    for (int si = 0; si < 4; si++) {  // same trip count in every lane: legs 1-3 see the torso slot as a null sphere
      const V3 crel = si == 0 ? mk(0.f, 0.f, 0.f) : (si == 1 ? r_tip : (si == 2 ? r_ank : K.rh));
      const float r = si == 0 ? (k == 0 ? ant::R_TORSO : -1e6f) : ant::R_CAPS;  // radius -1e6: every distance is huge
      const float body = si == 1 ? 2.f : (si == 2 ? 1.f : 0.f);
      const V3 c = s.O + crel;
      const float dg = c.z - P.gz - r;
      if (dg < P.margin) {
        if (si == 1 || si == 2) feet_ground = 1;
        nC = add_cand(cands, lane, nC, crel, r, mk(0.f, 0.f, 1.f), dg, body);
      }
      // everything but the ground is rare: the tests run out of line (off the sub-step loop's instruction footprint, which
      // sits at the edge of the 32 KB L1.5 instruction cache), behind cheap inline culls
      if (P.has_walls && fminf(P.wx - fabsf(c.x), P.wy - fabsf(c.y)) - r < P.margin)
        nC = sphere_vs_walls(c, crel, r, body, P.wx, P.wy, P.margin, cands, lane, nC);
      if (!ITEMS && P.has_box) {  // (AntGather has no box)
        const float reach = r + P.margin;
        const float ax = fmaxf(fmaxf(P.blo[0] - c.x, c.x - P.bhi[0]), 0.f), ay = fmaxf(fmaxf(P.blo[1] - c.y, c.y - P.bhi[1]), 0.f),
                    az = fmaxf(fmaxf(P.blo[2] - c.z, c.z - P.bhi[2]), 0.f);
        if (reach > 0.f && ax * ax + ay * ay + az * az < reach * reach * 1.00001f)
          nC = sphere_vs_aabb(c, crel, r, body, P.blo[0], P.blo[1], P.blo[2], P.bhi[0], P.bhi[1], P.bhi[2], P.margin, cands, lane, nC) & 0xff;
      }
      if (ITEMS) {
        for (unsigned m = imask; m; m &= m - 1) {  // candidate cubes in index order, after ground / walls like the oracle
          const int gi = __ffs(m) - 1;
          const float bx = ixy[2 * gi], by = ixy[2 * gi + 1], hh = P.item_half, rr = hh + r + P.margin;
          if (fabsf(c.x - bx) > rr || fabsf(c.y - by) > rr) continue;
          nC = sphere_vs_aabb(c, crel, r, body + 4.f /* friction class of the cubes */, bx - hh, by - hh, P.item_z - hh, bx + hh, by + hh,
                              P.item_z + hh, P.margin, cands, lane, nC);
          if (count_touch && (nC & 0x100)) HRL_TOUCH_ADD(itouch, gi);
          nC &= 0xff;
        }
      }
    }
    if (ITEMS && imask)  // cylinder part of the leg's capsules vs the cubes within reach: after the spheres, like the oracle
      nC = min(nC + capsules_vs_cubes(s.O, K.rh, r_ank, r_tip, imask, ixy, P.item_half, P.item_z, P.margin, cands + lane, 32, nC, itouch, count_touch),
               HRL_MAXC);
    // Maze box: the cylinder part of this leg's three capsules against the box's four vertical edges (the end-spheres
    // above only cover contacts at a capsule end; for a segment outside a convex rectangle the other closest pair is
    // (rectangle corner, segment interior); planar because the legs never reach the box's top).  Same order as the
    // oracle: foot, aux, leg capsule x corners (lo,lo) (hi,lo) (lo,hi) (hi,hi).  Culled per env on the torso distance.
    if (!ITEMS && P.has_box) {
      const float reach = 1.1314f + ant::R_CAPS + P.margin;  // |r_tip| <= 0.8 sqrt 2
      const float nx = fmaxf(fmaxf(P.blo[0] - s.O.x, s.O.x - P.bhi[0]), 0.f), ny = fmaxf(fmaxf(P.blo[1] - s.O.y, s.O.y - P.bhi[1]), 0.f);
      if (nx * nx + ny * ny < reach * reach)  // rare: out of line, off the sub-step loop's instruction footprint
        nC = capsules_vs_box_edges(s.O, K.rh, r_ank, r_tip, P.blo[0], P.blo[1], P.blo[2], P.bhi[0], P.bhi[1], P.bhi[2], P.margin, cands, lane, nC);
    }
  }

  if (dbg) {  // debugging builds (HRL_DEBUG_CONTACTS): this leg's candidate list
    dbg[0] = (float)nC;
    for (int c = 0; c < nC; c++)
      for (int f = 0; f < HRL_CAND_F; f++) dbg[1 + c * HRL_CAND_F + f] = CAND(c, f);
  }
  // ---------------- smooth dynamics: bias forces, leg elimination, base Schur complement ----------
  LegDyn D;
  float ub[6], u1, u2;  // generalized velocity after the unconstrained update
  {
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
    // F = m a - f_ext, N = I alpha + w x I w - tau_ext; f_ext = gravity + Bullet link damping
    const V3 F_a = ant::M_SHORT * (a_ac + gz) + (P.kl * ant::M_SHORT * (1.f + norm(v_ac))) * v_ac;
    const V3 N_a = axisym(ant::IX_SHORT, ant::IZ_SHORT, K.ez, al1) + cross(w1, Iw1) + (P.ka * (1.f + norm(w1))) * Iw1;
    const V3 F_f = ant::M_LONG * (a_fc + gz) + (P.kl * ant::M_LONG * (1.f + norm(v_fc))) * v_fc;
    const V3 N_f = axisym(ant::IX_LONG, ant::IZ_LONG, K.zf, al2) + cross(w2, Iw2) + (P.ka * (1.f + norm(w2))) * Iw2;
    const V3 F_l = (P.kl * ant::M_SHORT * (1.f + norm(v_lc))) * v_lc;  // fixed leg link: damping only
    const V3 N_l = (P.ka * (1.f + norm(w))) * Iwl;
    const float cb2 = dot(a2, N_f + cross(K.r2, F_f));
    const float cb1 = dot(a1, N_a + cross(K.r1, F_a) + N_f + cross(rf1, F_f));
    V3 cF = F_a + F_f + F_l;
    V3 cT = N_a + cross(r_ac, F_a) + N_f + cross(r_fc, F_f) + N_l + cross(0.5f * K.rh, F_l);

    // leg mass-matrix blocks
    const V3 lam2 = cross(a2, K.r2), lam1a = cross(a1, K.r1), lam1f = cross(a1, rf1);
    const V3 If_a1 = axisym(ant::IX_LONG, ant::IZ_LONG, K.zf, a1);
    const V3 If_a2 = ant::IX_LONG * a2;  // zf _|_ a2
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
    // articulated inertia of this leg about O: rigid6(aux) + rigid6(foot) - G Mll^-1 G^T
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
      // upper-right block m [d]x
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
    {  // + torso with its rigid leg capsules (COM at O)
      const float dzc = ant::IZ_COMP - ant::IX_COMP;
      const float za[3] = {K.ez.x, K.ez.y, K.ez.z};
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i; j < 3; j++) S[sym6(i, j)] += (i == j ? ant::IX_COMP : 0.f) + dzc * za[i] * za[j];
      S[sym6(3, 3)] += ant::M_COMP; S[sym6(4, 4)] += ant::M_COMP; S[sym6(5, 5)] += ant::M_COMP;
    }
    // Cholesky S = L L^T (in place, lower), then L^-1
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
    // generalized accelerations
    const float f1 = tau1 - cb1, f2 = tau2 - cb2;
    const float y1 = D.mi11 * f1 + D.mi12 * f2, y2 = D.mi12 * f1 + D.mi22 * f2;
    float fb[6] = {-cT.x, -cT.y, -cT.z, -cF.x, -cF.y, -cF.z};
#pragma unroll
    for (int i = 0; i < 6; i++) fb[i] = gsum(fb[i] - (D.G1[i] * y1 + D.G2[i] * y2));
    {  // torso composite: w x I w, gravity, damping of the torso sphere link
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

  // ---------------- constraint rows: counts, visit positions, whitened rows ----------------
  const float inv_h = P.inv_h;
  constexpr int ENV_F4 = COMPACT ? HRL_ENV_F4_COMPACT : HRL_ENV_F4, ROWS_FLOATS = HRL_EPW * ENV_F4 * 4;
  float4* __restrict__ rb = reinterpret_cast<float4*>(rows) + es * ENV_F4;  // es: this env's slot in the warp
  float4* __restrict__ cl = reinterpret_cast<float4*>(rows + ROWS_FLOATS + es * HRL_CL_STRIDE);  // contact slots
  float* __restrict__ lamL = rows + ROWS_FLOATS + HRL_CL_FLOATS_PER_WARP + es * HRL_LAML_STRIDE;     // limit impulses
  const float pl1 = s.q1 - ant::HIP_LO, ph1 = ant::HIP_HI - s.q1;
  const float pl2 = s.q2 - lc.lo2, ph2 = lc.hi2 - s.q2;
  // Bullet creates a joint-limit row iff the joint is at or beyond the limit (lo < hi: at most one per joint)
  const bool lim1 = (pl1 <= 0.f) || (ph1 <= 0.f), lim2 = (pl2 <= 0.f) || (ph2 <= 0.f);
  const int nL = (int)lim1 + (int)lim2;
  int incl = nL | (nC << 8);  // inclusive scan over the 4 legs of the env -> Bullet row order
  {
    int t = __shfl_up_sync(HRL_FULL_MASK, incl, 1, 4);
    if (k >= 1) incl += t;
    t = __shfl_up_sync(HRL_FULL_MASK, incl, 2, 4);
    if (k >= 2) incl += t;
  }
  const int tot = __shfl_sync(HRL_FULL_MASK, incl, 3, 4);
  const int offL = (incl & 0xff) - nL, offC = (incl >> 8) - nC, NL = tot & 0xff, NC = tot >> 8;
  const int maxNL = __reduce_max_sync(HRL_FULL_MASK, NL), maxNC = __reduce_max_sync(HRL_FULL_MASK, NC);
  // warp-uniform choice of the sweep: Delassus space (static slots, rows transposed) or velocity space (any row count)
  const bool fast = !COMPACT && HRL_DELASSUS && maxNC <= HRL_DS_MAXC;
  if (!COMPACT && HRL_DELASSUS) {
    if (fast) {  // clear the transposed rows, 1 / diag, rhs, leg of all 20 slots: idle slots and the other legs' components are 0
#pragma unroll
      for (int i = 0; i < HRL_DS_ZERO_F4 / 4; i++) rb[4 * i + k] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {     // the Delassus buffers of an earlier sub-step overlap the two all-zero rows of the idle visits
      rb[HRL_ROW_ZERO * 4 + 2 * k] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[HRL_ROW_ZERO * 4 + 2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
  }
  {
    const float zero6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int jj = 0; jj < 2; jj++) {
      if (!(jj ? lim2 : lim1)) continue;
      const float pl = jj ? pl2 : pl1, ph = jj ? ph2 : ph1;
      const bool lo = pl <= 0.f;
      const float sg = lo ? 1.f : -1.f, pen = lo ? pl : ph;
      const int pos = offL + (jj ? (int)lim1 : 0);
      HRL_CHECK(pos >= 0 && pos < 8);
      lamL[pos] = 0.f;
      emit_row<COMPACT>(rb, pos, k, D, zero6, jj ? 0.f : sg, jj ? sg : 0.f, ub, u1, u2, pen, P.erp_l, inv_h, true, fast, pos);
    }
  }
  for (int c = 0; c < nC; c++) {
    const V3 Pr = mk(CAND(c, 0), CAND(c, 1), CAND(c, 2));
    const V3 n = mk(CAND(c, 3), CAND(c, 4), CAND(c, 5));
    const float dist = CAND(c, 6), bodyc = CAND(c, 7);
    const bool cube = ITEMS && bodyc >= 4.f;
    const float body = cube ? bodyc - 4.f : bodyc;
    V3 t1, t2;
    plane_space(n, t1, t2);
    const V3 Ph = Pr - K.rh, Pa = Pr - r_ank;
    const int ci = offC + c;
    HRL_CHECK(ci >= 0 && ci < HRL_NSLOT && c < HRL_MAXC);
    cl[ci] = make_float4(0.f, 0.f, 0.f, cube ? P.mu_item : P.mu);
#pragma unroll 1
    for (int di = 0; di < 3; di++) {
      const V3 d = di == 0 ? n : (di == 1 ? t1 : t2);
      const V3 jt = cross(Pr, d);
      const float JB[6] = {jt.x, jt.y, jt.z, d.x, d.y, d.z};
      const float j1 = body >= 1.f ? dot(a1, cross(Ph, d)) : 0.f;
      const float j2 = body >= 2.f ? dot(a2, cross(Pa, d)) : 0.f;
      const int pos = di == 0 ? HRL_ROW_NRM_LAST - ci : HRL_ROW_FRI0 + 2 * ci + (di - 1);
      emit_row<COMPACT>(rb, pos, k, D, JB, j1, j2, ub, u1, u2, dist, P.erp_c, inv_h, di == 0, fast,
                        8 + 4 * ci + di);
    }
  }
  stat_contacts += nC; stat_limits += nL;
  __syncwarp();

  // ---------------- projected Gauss-Seidel, Bullet row order ----------------
  // Every lane of the env evaluates every row (broadcast shared-memory reads, no shuffle on the
  // dependency chain).  A single warp issues at most one instruction per ~2 cycles, so the sweep is
  // bound by its instruction count: packed FFMA2 arithmetic, rows prefetched one visit ahead into
  // ping-pong registers, trip counts = maxima over the envs of the warp, idle visits on a zero row.
  if (trips) *trips += maxNL + (maxNC << 16);  // profiling builds (HRL_WARP_TIMES): the warp's solver trip counts
  float dvp[6], gp0, gp1;  // solver result in the transformed velocity: base part, this leg's part
  if (!COMPACT && HRL_DELASSUS && fast) {
    // ---------------- Delassus-space sweep (layout and derivation at the top of this file) ----------------
    float* __restrict__ f = reinterpret_cast<float*>(rb);
    const unsigned gmask = (maxNL > 0 ? 1u : 0u) | (maxNL > 4 ? 2u : 0u) | (maxNC > 0 ? 4u : 0u) | (maxNC > 1 ? 8u : 0u) |
                           (maxNC > 2 ? 16u : 0u) | (maxNC > 3 ? 32u : 0u);
    ds_gram(f, k, gmask);
    DsState st;
    {  // lambda = 0, dv' = 0: t = rhs'
      const float4 q0 = *reinterpret_cast<const float4*>(f + HRL_DS_RHS), q1 = *reinterpret_cast<const float4*>(f + HRL_DS_RHS + 4);
      st.tL[0] = make_float2(q0.x, q0.y); st.tL[1] = make_float2(q0.z, q0.w); st.tL[2] = make_float2(q1.x, q1.y); st.tL[3] = make_float2(q1.z, q1.w);
      st.tC = *reinterpret_cast<const float4*>(f + HRL_DS_RHS + 8 + 4 * k);
#pragma unroll
      for (int i = 0; i < 4; i++) st.lL[i] = make_float2(0.f, 0.f);
      st.lC = make_float4(0.f, 0.f, 0.f, 0.f);
      st.mu = cl[k].w;
    }
    __syncwarp();  // What complete
    const float* __restrict__ W = f + HRL_DS_W;
#define HRL_DS_L(S) ds_limit<S>(st, W, k, P.max_imp)
#pragma unroll 1
    for (int it = 0; it < P.iters; it++) {
      // (1) joint-limit rows, backwards on even iterations (idle positions between an env's rows are no-ops: t = lambda = 0)
      if (it & 1) {
        if (maxNL > 0) { HRL_DS_L(0); HRL_DS_L(1); }
        if (maxNL > 2) { HRL_DS_L(2); HRL_DS_L(3); }
        if (maxNL > 4) { HRL_DS_L(4); HRL_DS_L(5); }
        if (maxNL > 6) { HRL_DS_L(6); HRL_DS_L(7); }
      } else {
        switch (maxNL) {
          case 8: HRL_DS_L(7);
          case 7: HRL_DS_L(6);
          case 6: HRL_DS_L(5);
          case 5: HRL_DS_L(4);
          case 4: HRL_DS_L(3);
          case 3: HRL_DS_L(2);
          case 2: HRL_DS_L(1);
          case 1: HRL_DS_L(0);
          default: break;
        }
      }
      // (2) contact normals, (3) friction pairs (idle contacts are no-ops - t = lambda = 0, What row 0 - so whole blocks
      // run without exits in between)
      if (maxNC > 0) {
        ds_normal<0>(st, W, k); ds_normal<1>(st, W, k);
        if (maxNC > 2) { ds_normal<2>(st, W, k); ds_normal<3>(st, W, k); }
        ds_pair<0>(st, W, k); ds_pair<1>(st, W, k);
        if (maxNC > 2) { ds_pair<2>(st, W, k); ds_pair<3>(st, W, k); }
      }
    }
#undef HRL_DS_L
    // dv' = sum_p lambda_p a_p: the 6 base components and the 2 of this lane's leg
    float2 lam[HRL_DS_P / 2];
#pragma unroll
    for (int i = 0; i < 4; i++) lam[i] = st.lL[i];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      lam[4 + 2 * c] = make_float2(__shfl_sync(HRL_FULL_MASK, st.lC.x, c, 4), __shfl_sync(HRL_FULL_MASK, st.lC.y, c, 4));
      lam[5 + 2 * c] = make_float2(__shfl_sync(HRL_FULL_MASK, st.lC.z, c, 4), 0.f);
    }
    float o[8];
#pragma unroll
    for (int mm = 0; mm < 8; mm++) {
      const float4* __restrict__ a = reinterpret_cast<const float4*>(f + HRL_DS_AT + (mm < 6 ? mm : 2 * k + mm) * HRL_DS_P);
      float2 s0 = make_float2(0.f, 0.f), s1 = s0;
#pragma unroll
      for (int g = 0; g < HRL_DS_P / 4; g++) {
        const float4 q = a[g];
        s0 = fma2(make_float2(q.x, q.y), lam[2 * g], s0);
        s1 = fma2(make_float2(q.z, q.w), lam[2 * g + 1], s1);
      }
      s0 = add2(s0, s1);
      o[mm] = s0.x + s0.y;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) dvp[i] = o[i];
    gp0 = o[6]; gp1 = o[7];
  } else {
  float2 dv[7];
#pragma unroll
  for (int i = 0; i < 7; i++) dv[i] = make_float2(0.f, 0.f);
#define HRL_LIM_ROW(t) (((t) < NL) ? lbase + lstep * (t) : HRL_ROW_ZERO)
#define HRL_LIM_LAM(t, r) (lamL + (((t) < NL) ? (r) : 8))
#define HRL_SLOT(t) (((t) < NC) ? (t) : HRL_NSLOT)
  HRL_CHECK(NL >= 0 && NL <= 8 && NC >= 0 && NC <= HRL_NSLOT && maxNL <= 8 && maxNC <= HRL_NSLOT);
  for (int it = 0; it < P.iters; it++) {
    // (1) joint-limit rows; Bullet walks the non-contact rows backwards on even iterations.
    // Two visits per trip (ping-pong row registers, next row prefetched), odd tail handled apart.
    if (maxNL > 0) {
      const int lbase = (it & 1) ? 0 : NL - 1, lstep = (it & 1) ? 1 : -1;
      int r0 = HRL_LIM_ROW(0), r1;
      float *p0 = HRL_LIM_LAM(0, r0), *p1;
      Row R0 = ld_row<COMPACT>(rb, r0), R1;
      float l0 = *p0, l1;
      int t = 0;
      for (; t + 1 < maxNL; t += 2) {
        r1 = HRL_LIM_ROW(t + 1); p1 = HRL_LIM_LAM(t + 1, r1); R1 = ld_row<COMPACT>(rb, r1); l1 = *p1;
#if HRL_LOOKAHEAD
        // look-ahead inside the trip: both rows are in registers, so the second visit's row velocity is taken against the
        // velocity BEFORE the first visit (off the chain, next to it) and corrected with the rows' Gram entry:
        // a1 . (dv + a0 dl0) = a1 . dv + (a0 . a1) dl0 - its dependent chain is fma -> fma -> max -> min -> sub
        const float g01 = dot14(R0.p, R1.p), d1 = dot14(R1.p, dv);
        const float dl0 = single_visit<true>(p0, dv, R0, l0, P.max_imp);
        r0 = HRL_LIM_ROW(t + 2); p0 = HRL_LIM_LAM(t + 2, r0); R0 = ld_row<COMPACT>(rb, r0); l0 = *p0;
        single_visit_pre<true>(p1, dv, R1, l1, P.max_imp, fmaf(g01, dl0, d1));
#else
        single_visit<true>(p0, dv, R0, l0, P.max_imp);
        r0 = HRL_LIM_ROW(t + 2); p0 = HRL_LIM_LAM(t + 2, r0); R0 = ld_row<COMPACT>(rb, r0); l0 = *p0;
        single_visit<true>(p1, dv, R1, l1, P.max_imp);
#endif
      }
      if (t < maxNL) single_visit<true>(p0, dv, R0, l0, P.max_imp);
    }
    if (maxNC > 0) {
      // (2) contact normals: slot c at row NRM_LAST - c, impulse in cl[c].z
      {
        int c0 = HRL_SLOT(0), c1;
        Row R0 = ld_row<COMPACT>(rb, HRL_ROW_NRM_LAST - c0), R1;
        float l0 = cl[c0].z, l1;
        int t = 0;
        for (; t + 1 < maxNC; t += 2) {
          c1 = HRL_SLOT(t + 1); R1 = ld_row<COMPACT>(rb, HRL_ROW_NRM_LAST - c1); l1 = cl[c1].z;
#if HRL_LOOKAHEAD
          const float g01 = dot14(R0.p, R1.p), d1 = dot14(R1.p, dv);
          const float dl0 = single_visit<false>(&cl[c0].z, dv, R0, l0, 0.f);
          c0 = HRL_SLOT(t + 2); R0 = ld_row<COMPACT>(rb, HRL_ROW_NRM_LAST - c0); l0 = cl[c0].z;
          single_visit_pre<false>(&cl[c1].z, dv, R1, l1, 0.f, fmaf(g01, dl0, d1));
#else
          single_visit<false>(&cl[c0].z, dv, R0, l0, 0.f);
          c0 = HRL_SLOT(t + 2); R0 = ld_row<COMPACT>(rb, HRL_ROW_NRM_LAST - c0); l0 = cl[c0].z;
          single_visit<false>(&cl[c1].z, dv, R1, l1, 0.f);
#endif
        }
        if (t < maxNC) single_visit<false>(&cl[c0].z, dv, R0, l0, 0.f);
      }
      // (3) friction pairs: slot c at rows FRI0 + 2c, + 1; (lambda_t1, lambda_t2, lambda_n, mu) = cl[c]
      {
        int c0 = HRL_SLOT(0), c1;
        Row A0 = ld_row<COMPACT>(rb, HRL_ROW_FRI0 + 2 * c0), B0 = ld_row<COMPACT>(rb, HRL_ROW_FRI0 + 2 * c0 + 1), A1, B1;
        float4 q0 = cl[c0], q1;
        int t = 0;
        for (; t + 1 < maxNC; t += 2) {
          c1 = HRL_SLOT(t + 1); A1 = ld_row<COMPACT>(rb, HRL_ROW_FRI0 + 2 * c1); B1 = ld_row<COMPACT>(rb, HRL_ROW_FRI0 + 2 * c1 + 1); q1 = cl[c1];
          pair_visit(cl + c0, dv, A0, B0, q0);
          c0 = HRL_SLOT(t + 2); A0 = ld_row<COMPACT>(rb, HRL_ROW_FRI0 + 2 * c0); B0 = ld_row<COMPACT>(rb, HRL_ROW_FRI0 + 2 * c0 + 1); q0 = cl[c0];
          pair_visit(cl + c1, dv, A1, B1, q1);
        }
        if (t < maxNC) pair_visit(cl + c0, dv, A0, B0, q0);
      }
    }
  }
#undef HRL_LIM_LAM
#undef HRL_SLOT
#undef HRL_LIM_ROW
#pragma unroll
  for (int j = 0; j < 6; j++) dvp[j] = (j & 1) ? dv[j >> 1].y : dv[j >> 1].x;
  const float2 gp = k == 0 ? dv[3] : (k == 1 ? dv[4] : (k == 2 ? dv[5] : dv[6]));
  gp0 = gp.x; gp1 = gp.y;
  }
  __syncwarp();

  // ---------------- back to physical velocities, clamp, integrate ----------------
  float dvb[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {  // dvb = L^-T dvb'
    float a = 0.f;
#pragma unroll
    for (int j = i; j < 6; j++) a = fmaf(D.Li[j * (j + 1) / 2 + i], dvp[j], a);
    dvb[i] = a;
  }
  const float g2 = gp1 * D.il22, g1 = (gp0 - D.l21 * g2) * D.il11;  // g = Ll^-T g'
  float dq1 = g1, dq2 = g2;
#pragma unroll
  for (int i = 0; i < 6; i++) { dq1 = fmaf(-D.K0[i], dvb[i], dq1); dq2 = fmaf(-D.K1[i], dvb[i], dq2); }
#pragma unroll
  for (int i = 0; i < 6; i++) ub[i] = clampf(ub[i] + dvb[i], P.vmax);
  u1 = clampf(u1 + dq1, P.vmax); u2 = clampf(u2 + dq2, P.vmax);
  s.w = mk(ub[0], ub[1], ub[2]); s.v = mk(ub[3], ub[4], ub[5]);
  s.qd1 = u1; s.qd2 = u2;
  s.O = s.O + P.h * s.v;
  s.q1 = fmaf(P.h, u1, s.q1); s.q2 = fmaf(P.h, u2, s.q2);
  {  // q <- exp(w h) * q  (Bullet pQuatUpdateFun with world-frame omega)
    // half angle y = |w| h / 2: exp = (w * sin(y)/|w|, cos(y)).  Bullet caps |w| h at pi/4, so y <= pi/8
    // and the even Taylor polynomials in y^2 below are exact to 2e-9 (sinc) / 2e-11 (cos): no sqrt,
    // no division, no branch.  |w| h > pi/4 (needs max_coord_vel > 110) takes the literal path.
    const float w2 = dot(s.w, s.w), y2 = 0.25f * P.h * P.h * w2;
    float kk, cw;
    if (y2 <= 0.1542126f) {  // (pi/8)^2
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
