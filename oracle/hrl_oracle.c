/*
 * hrl_oracle.c - CPU restatement of the hrl_pybullet_envs env-step.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (hrl_pybullet_envs_b200/) never does.
 *
 * PARITY STATUS
 *   - task layer (Gather sensor, pickups/respawn rule, wall lidar, maze goal, Flagrun target
 *     logic, reward/done): PINNED against golden vectors produced by executing the reference's
 *     own Python (tests/golden/make_golden.py -> the tests/golden npz files).
 *   - physics: "PARITY UNPINNED".  The arithmetic lives in pybullet (Bullet C++,
 *     requirements.txt:1 `pybullet>=3.0.0`, unpinned, not vendored, not installable here).  This
 *     file restates the published algorithms Bullet's btMultiBody pipeline implements
 *     (Featherstone ABA in link coordinates, unilateral joint-limit and contact rows, projected
 *     Gauss-Seidel with 5 iterations, symplectic Euler) with the constants recalled in
 *     SURVEY.md App. A.  It is written in a deliberately different formulation from the CUDA
 *     path (13 separate links, link-local spatial algebra, O(n) ABA + ABA impulse response)
 *     so that agreement between the two is a real cross-check.
 *
 * Citations `file:line` are relative to /root/reference/hrl_pybullet_envs/.
 * Build: see oracle/Makefile (double: libhrl_oracle.so; float: libhrl_oracle_f32.so).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hrl_b200.h"

#if defined(HRLO_COUNT) /* C++ build with a counting `real` (oracle/flop_counter.hpp, tools/count_oracle_flops.py) */
typedef creal real;
#define R_SQRT c_sqrt
#define R_SIN c_sin
#define R_COS c_cos
#define R_ATAN2 c_atan2
#define R_ASIN c_asin
#define R_FABS c_fabs
#define R_FLOOR c_floor
#elif defined(HRLO_F32)
typedef float real;
#define R_SQRT sqrtf
#define R_SIN sinf
#define R_COS cosf
#define R_ATAN2 atan2f
#define R_ASIN asinf
#define R_FABS fabsf
#define R_FLOOR floorf
#else
typedef double real;
#define R_SQRT sqrt
#define R_SIN sin
#define R_COS cos
#define R_ATAN2 atan2
#define R_ASIN asin
#define R_FABS fabs
#define R_FLOOR floor
#endif

#define PI_D 3.14159265358979323846

/* ======================================================================================
 * Counter-based RNG shared (as a specification) with the CUDA path: Philox4x32-10
 * (Salmon et al., SC'11).  The reference uses MT19937 RandomState streams
 * (gather_scene.py:31,58; ant_maze_bullet_env.py:48,110; ant_flagrun_env.py:39) which are not
 * reproduced: distributional parity only (SURVEY.md hard part 6).
 * ====================================================================================== */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* RNG addressing (no sequential per-env counters, so the GPU lanes can draw independently):
 *   joint noise      : stream JOINT,      draw = episode,            sub = 0/1 (joints 0-3 / 4-7)
 *   item respawn     : stream ITEM,       draw = env total steps,    sub = item*64 + attempt
 *   item reset       : stream ITEM_RESET, draw = episode,            sub = item*64 + attempt
 *   maze goal        : stream GOAL,       draw = episode,            sub = 0
 *   flagrun goal j   : stream FLAG (key flag_seed, env 0), draw = attempt, sub = episode*128 + j */
enum { STREAM_JOINT = 0, STREAM_ITEM = 1, STREAM_GOAL = 2, STREAM_FLAG = 3, STREAM_ITEM_RESET = 4, STREAM_FLAG_CLOSE = 5 };
#define MAX_PLACE_ATTEMPTS 16
#define MAX_CLOSE_ATTEMPTS 64
/* 4 uniforms in [0,1) with 24 bits each (exact in float and double) */
static void rng_u4(uint64_t seed, uint32_t env, uint32_t stream, uint32_t draw, uint32_t sub, real u[4]) {
  uint32_t o[4];
  philox4x32_10(draw, env, stream, sub, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  for (int i = 0; i < 4; i++) u[i] = (real)(o[i] >> 8) * (real)(1.0 / 16777216.0);
}

/* ======================================================================================
 * small linear algebra
 * ====================================================================================== */
typedef struct { real v[3]; } v3;
typedef struct { real m[3][3]; } m3;
typedef struct { real v[6]; } sv;     /* spatial vector [angular; linear] */
typedef struct { real m[6][6]; } sm;  /* spatial matrix */

static v3 V3(real x, real y, real z) { v3 r = {{x, y, z}}; return r; }
static v3 vadd(v3 a, v3 b) { return V3(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }
static v3 vsub(v3 a, v3 b) { return V3(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }
static v3 vscale(v3 a, real s) { return V3(a.v[0] * s, a.v[1] * s, a.v[2] * s); }
static real vdot(v3 a, v3 b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
static v3 vcross(v3 a, v3 b) {
  return V3(a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]);
}
static real vnorm(v3 a) { return R_SQRT(vdot(a, a)); }
static v3 mmulv(const m3* A, v3 x) {
  v3 r;
  for (int i = 0; i < 3; i++) r.v[i] = A->m[i][0] * x.v[0] + A->m[i][1] * x.v[1] + A->m[i][2] * x.v[2];
  return r;
}
static v3 mtmulv(const m3* A, v3 x) {
  v3 r;
  for (int i = 0; i < 3; i++) r.v[i] = A->m[0][i] * x.v[0] + A->m[1][i] * x.v[1] + A->m[2][i] * x.v[2];
  return r;
}
static m3 mmul(const m3* A, const m3* B) {
  m3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = A->m[i][0] * B->m[0][j] + A->m[i][1] * B->m[1][j] + A->m[i][2] * B->m[2][j];
  return C;
}
static m3 mtrans(const m3* A) {
  m3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = A->m[j][i];
  return C;
}
static m3 skew(v3 a) {
  m3 S = {{{0, -a.v[2], a.v[1]}, {a.v[2], 0, -a.v[0]}, {-a.v[1], a.v[0], 0}}};
  return S;
}
/* rotation by angle q about unit axis a (Rodrigues) */
static m3 rot_axis(v3 a, real q) {
  real c = R_COS(q), s = R_SIN(q), t = 1 - c;
  m3 R = {{{t * a.v[0] * a.v[0] + c, t * a.v[0] * a.v[1] - s * a.v[2], t * a.v[0] * a.v[2] + s * a.v[1]},
           {t * a.v[0] * a.v[1] + s * a.v[2], t * a.v[1] * a.v[1] + c, t * a.v[1] * a.v[2] - s * a.v[0]},
           {t * a.v[0] * a.v[2] - s * a.v[1], t * a.v[1] * a.v[2] + s * a.v[0], t * a.v[2] * a.v[2] + c}}};
  return R;
}
static m3 quat_to_m3(const real q[4]) { /* x,y,z,w */
  real x = q[0], y = q[1], z = q[2], w = q[3];
  m3 R = {{{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)},
           {2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)},
           {2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}}};
  return R;
}
static sv sv_zero(void) { sv r; memset(&r, 0, sizeof r); return r; }
static sv sv_make(v3 a, v3 l) { sv r = {{a.v[0], a.v[1], a.v[2], l.v[0], l.v[1], l.v[2]}}; return r; }
static v3 sv_ang(sv s) { return V3(s.v[0], s.v[1], s.v[2]); }
static v3 sv_lin(sv s) { return V3(s.v[3], s.v[4], s.v[5]); }
static sv sv_add(sv a, sv b) { sv r; for (int i = 0; i < 6; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
static sv sv_sub(sv a, sv b) { sv r; for (int i = 0; i < 6; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
static sv sv_scale(sv a, real s) { sv r; for (int i = 0; i < 6; i++) r.v[i] = a.v[i] * s; return r; }
static real sv_dot(sv a, sv b) { real s = 0; for (int i = 0; i < 6; i++) s += a.v[i] * b.v[i]; return s; }
static sv sm_mulv(const sm* A, sv x) {
  sv r;
  for (int i = 0; i < 6; i++) { real s = 0; for (int j = 0; j < 6; j++) s += A->m[i][j] * x.v[j]; r.v[i] = s; }
  return r;
}
static sv sm_tmulv(const sm* A, sv x) {
  sv r;
  for (int i = 0; i < 6; i++) { real s = 0; for (int j = 0; j < 6; j++) s += A->m[j][i] * x.v[j]; r.v[i] = s; }
  return r;
}
/* motion cross product v x m, force cross product v x* f (Featherstone RBDA eq. 2.31/2.32) */
static sv crm(sv v, sv m) {
  v3 w = sv_ang(v), vl = sv_lin(v), ma = sv_ang(m), ml = sv_lin(m);
  return sv_make(vcross(w, ma), vadd(vcross(w, ml), vcross(vl, ma)));
}
static sv crf(sv v, sv f) {
  v3 w = sv_ang(v), vl = sv_lin(v), n = sv_ang(f), fl = sv_lin(f);
  return sv_make(vadd(vcross(w, n), vcross(vl, fl)), vcross(w, fl));
}
/* Cholesky of SPD 6x6 (lower), solve */
static int chol6(const sm* A, sm* L) {
  memset(L, 0, sizeof *L);
  for (int j = 0; j < 6; j++) {
    real d = A->m[j][j];
    for (int k = 0; k < j; k++) d -= L->m[j][k] * L->m[j][k];
    if (!(d > 0)) return -1;
    L->m[j][j] = R_SQRT(d);
    for (int i = j + 1; i < 6; i++) {
      real s = A->m[i][j];
      for (int k = 0; k < j; k++) s -= L->m[i][k] * L->m[j][k];
      L->m[i][j] = s / L->m[j][j];
    }
  }
  return 0;
}
static sv chol6_solve(const sm* L, sv b) {
  sv y, x;
  for (int i = 0; i < 6; i++) { real s = b.v[i]; for (int k = 0; k < i; k++) s -= L->m[i][k] * y.v[k]; y.v[i] = s / L->m[i][i]; }
  for (int i = 5; i >= 0; i--) { real s = y.v[i]; for (int k = i + 1; k < 6; k++) s -= L->m[k][i] * x.v[k]; x.v[i] = s / L->m[i][i]; }
  return x;
}

/* ======================================================================================
 * Ant model (assets/ant.xml).  13 links: torso, and per leg k the
 * fixed `*_leg` capsule link, `aux_k` (hinge hip_k) and the foot link (hinge ankle_k).
 * Masses = 1000 kg/m^3 x capsule/sphere volume; inertias = Bullet's compound-shape AABB rule
 * (SURVEY.md A.3 "Import") - recalled, unverified.
 * ====================================================================================== */
#define NLINK 13
#define NDOF 14 /* 6 base + 8 hinges; generalized velocity = [w_world, v_world(origin), qd] */
static const int LEG_SX[4] = {+1, -1, -1, +1}; /* ant.xml:15,26,37,48 */
static const int LEG_SY[4] = {+1, +1, -1, -1};
#define ANT_R_TORSO 0.25  /* ant.xml:13 */
#define ANT_R_CAPS 0.08   /* ant.xml:16 */
#define ANT_M_TORSO 65.44984694978736      /* 1000*(4/3)pi*0.25^3 */
#define ANT_I_TORSO 2.7270769562411402     /* m/12*(0.5^2+0.5^2) (AABB rule) */
#define ANT_M_SHORT 7.831583314284915      /* 1000*(pi r^2 L + 4/3 pi r^3), L = 0.2*sqrt(2) */
#define ANT_IXX_SHORT 0.10128847753141823  /* m/12*(0.36^2+0.16^2) */
#define ANT_IZZ_SHORT 0.16916219958855416  /* m/12*(0.36^2+0.36^2) */
#define ANT_M_LONG 13.51850726010076       /* L = 0.4*sqrt(2) */
#define ANT_IXX_LONG 0.3821231385521815    /* m/12*(0.56^2+0.16^2) */
#define ANT_IZZ_LONG 0.706567312794599     /* m/12*(0.56^2+0.56^2) */
#define HIP_LO (-0.6981317007977318)  /* ant.xml:18 range -40..40 deg */
#define HIP_HI (0.6981317007977318)
#define ANK_LO (0.5235987755982988)   /* ant.xml:21 range 30..100 deg (legs 1,4); negated for legs 2,3 */
#define ANK_HI (1.7453292519943295)

typedef struct {
  int parent, jtype, dof; /* jtype 0 fixed, 1 revolute; dof index 0..7 */
  v3 axis, r, com;        /* joint axis (link coords), origin in parent coords, COM in link coords */
  real mass, ixx, izz;    /* inertia diag (ixx, ixx, izz) about COM in link axes */
} link_t;

typedef struct { int link; v3 local; real radius; int group; int foot; } sphere_t;
#define NSPHERE 13

typedef struct {
  link_t L[NLINK];
  real lo[8], hi[8];
  sphere_t S[NSPHERE];
} ant_model;

static void ant_model_init(ant_model* M) {
  const real is2 = (real)0.70710678118654752440;
  memset(M, 0, sizeof *M);
  M->L[0].parent = -1; M->L[0].jtype = 0; M->L[0].dof = -1;
  M->L[0].mass = (real)ANT_M_TORSO; M->L[0].ixx = (real)ANT_I_TORSO; M->L[0].izz = (real)ANT_I_TORSO;
  int ns = 0;
  M->S[ns].link = 0; M->S[ns].local = V3(0, 0, 0); M->S[ns].radius = (real)ANT_R_TORSO; M->S[ns].group = 0; M->S[ns].foot = -1; ns++;
  for (int k = 0; k < 4; k++) {
    real sx = (real)LEG_SX[k], sy = (real)LEG_SY[k];
    link_t* leg = &M->L[1 + 3 * k]; link_t* aux = &M->L[2 + 3 * k]; link_t* foot = &M->L[3 + 3 * k];
    leg->parent = 0; leg->jtype = 0; leg->dof = -1; leg->r = V3(0, 0, 0);
    leg->com = V3((real)0.1 * sx, (real)0.1 * sy, 0); leg->mass = (real)ANT_M_SHORT; leg->ixx = (real)ANT_IXX_SHORT; leg->izz = (real)ANT_IZZ_SHORT;
    aux->parent = 1 + 3 * k; aux->jtype = 1; aux->dof = 2 * k; aux->axis = V3(0, 0, 1); /* ant.xml:18 */
    aux->r = V3((real)0.2 * sx, (real)0.2 * sy, 0); aux->com = V3((real)0.1 * sx, (real)0.1 * sy, 0);
    aux->mass = (real)ANT_M_SHORT; aux->ixx = (real)ANT_IXX_SHORT; aux->izz = (real)ANT_IZZ_SHORT;
    foot->parent = 2 + 3 * k; foot->jtype = 1; foot->dof = 2 * k + 1;
    /* ant.xml:21,32,43,54: ankle_1/3 axis (-1,1,0), ankle_2/4 axis (1,1,0), normalised */
    foot->axis = (k == 0 || k == 2) ? V3(-is2, is2, 0) : V3(is2, is2, 0);
    foot->r = V3((real)0.2 * sx, (real)0.2 * sy, 0); foot->com = V3((real)0.2 * sx, (real)0.2 * sy, 0);
    foot->mass = (real)ANT_M_LONG; foot->ixx = (real)ANT_IXX_LONG; foot->izz = (real)ANT_IZZ_LONG;
    M->lo[2 * k] = (real)HIP_LO; M->hi[2 * k] = (real)HIP_HI;
    if (k == 0 || k == 3) { M->lo[2 * k + 1] = (real)ANK_LO; M->hi[2 * k + 1] = (real)ANK_HI; }
    else { M->lo[2 * k + 1] = (real)-ANK_HI; M->hi[2 * k + 1] = (real)-ANK_LO; }
    /* collision spheres = capsule end-spheres, de-duplicated at shared joints:
       T = foot tip (foot link), A = ankle point (aux link far end), H = hip point (leg link far end) */
    M->S[ns].link = 3 + 3 * k; M->S[ns].local = V3((real)0.4 * sx, (real)0.4 * sy, 0); M->S[ns].radius = (real)ANT_R_CAPS; M->S[ns].group = k; M->S[ns].foot = k; ns++;
    M->S[ns].link = 2 + 3 * k; M->S[ns].local = V3((real)0.2 * sx, (real)0.2 * sy, 0); M->S[ns].radius = (real)ANT_R_CAPS; M->S[ns].group = k; M->S[ns].foot = k; ns++;
    M->S[ns].link = 1 + 3 * k; M->S[ns].local = V3((real)0.2 * sx, (real)0.2 * sy, 0); M->S[ns].radius = (real)ANT_R_CAPS; M->S[ns].group = k; M->S[ns].foot = -1; ns++;
  }
}

/* ======================================================================================
 * per-env state
 * ====================================================================================== */
typedef struct {
  real pos[3], quat[4], vel[3], ang[3], q[8], qd[8];
  real initial_z, potential, target[2], wtd, feet[4];
  real ret, ret_sum; /* episode-return accumulators (HRL_SF_RETURN, HRL_SF_RETURN_SUM) */
  real items[HRL_MAX_ITEMS][2];
  int32_t t, episode, steps_total, goals_left, since, rewarded, goal_gen;
} env_state;

struct hrlo_env {
  hrl_config cfg;
  ant_model model;
  env_state* s;
  /* instrumentation: flop-model inputs */
  double n_contacts, n_limit_rows, n_substeps, n_capsule_contacts; /* the last: contacts of a capsule's cylinder part */
  /* golden-vector replay (tests only): goal list instead of the Philox stream, stub robot */
  const double* replay_goals; /* [flag_max_targets][2], popped from the end like the reference */
  int replay_stub_robot;      /* next_target(): calc_potential() = -1, calc_state() leaves wtd alone */
};
typedef struct hrlo_env hrlo_env;

/* ======================================================================================
 * kinematics + ABA cache
 * ====================================================================================== */
typedef struct {
  m3 Rw[NLINK];   /* link -> world */
  v3 ow[NLINK];   /* link origin, world */
  v3 comw[NLINK]; /* link COM, world */
  sm X[NLINK];    /* motion transform parent -> link coords */
  sv v[NLINK], c[NLINK];
  sm IA[NLINK];
  sv U[NLINK];
  real Dinv[NLINK];
  sm L0; /* Cholesky of base articulated inertia */
} kin_t;

static sm spatial_inertia(const link_t* l) {
  sm I; memset(&I, 0, sizeof I);
  m3 C = skew(l->com);
  m3 CC = mmul(&C, &C);
  real Ic[3] = {l->ixx, l->ixx, l->izz};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      I.m[i][j] = (i == j ? Ic[i] : 0) - l->mass * CC.m[i][j];
      I.m[i][3 + j] = l->mass * C.m[i][j];
      I.m[3 + i][j] = -l->mass * C.m[i][j];
      I.m[3 + i][3 + j] = (i == j) ? l->mass : 0;
    }
  return I;
}

static void forward_kinematics(const ant_model* M, const env_state* s, kin_t* K) {
  K->Rw[0] = quat_to_m3(s->quat);
  K->ow[0] = V3(s->pos[0], s->pos[1], s->pos[2]);
  v3 wl = mtmulv(&K->Rw[0], V3(s->ang[0], s->ang[1], s->ang[2]));
  v3 vl = mtmulv(&K->Rw[0], V3(s->vel[0], s->vel[1], s->vel[2]));
  K->v[0] = sv_make(wl, vl);
  K->c[0] = sv_zero();
  K->comw[0] = vadd(K->ow[0], mmulv(&K->Rw[0], M->L[0].com));
  for (int i = 1; i < NLINK; i++) {
    const link_t* l = &M->L[i];
    int p = l->parent;
    m3 Rpc; /* link coords -> parent coords */
    real qd = 0;
    if (l->jtype == 1) { Rpc = rot_axis(l->axis, s->q[l->dof]); qd = s->qd[l->dof]; }
    else { memset(&Rpc, 0, sizeof Rpc); Rpc.m[0][0] = Rpc.m[1][1] = Rpc.m[2][2] = 1; }
    m3 E = mtrans(&Rpc);
    K->Rw[i] = mmul(&K->Rw[p], &Rpc);
    K->ow[i] = vadd(K->ow[p], mmulv(&K->Rw[p], l->r));
    K->comw[i] = vadd(K->ow[i], mmulv(&K->Rw[i], l->com));
    /* X = [[E, 0], [-E r~, E]] */
    m3 rx = skew(l->r); m3 Erx = mmul(&E, &rx);
    memset(&K->X[i], 0, sizeof(sm));
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        K->X[i].m[a][b] = E.m[a][b]; K->X[i].m[3 + a][3 + b] = E.m[a][b]; K->X[i].m[3 + a][b] = -Erx.m[a][b];
      }
    sv vJ = sv_make(vscale(l->axis, qd), V3(0, 0, 0));
    K->v[i] = sv_add(sm_mulv(&K->X[i], K->v[p]), vJ);
    K->c[i] = crm(K->v[i], vJ);
  }
}

/* external spatial force on link i in link coords: gravity + Bullet's per-link damping
 * m v (k + k|v|), I w (k + k|w|)   */
static sv link_ext_force(const hrl_config* cfg, const link_t* l, const m3* Rw, sv v) {
  v3 w = sv_ang(v);
  v3 vc = vadd(sv_lin(v), vcross(w, l->com));
  v3 fg = mtmulv(Rw, V3(0, 0, -l->mass * (real)cfg->gravity));
  real kl = (real)cfg->lin_damping, ka = (real)cfg->ang_damping;
  v3 fd = vscale(vc, -l->mass * (kl + kl * vnorm(vc)));
  v3 Iw = V3(l->ixx * w.v[0], l->ixx * w.v[1], l->izz * w.v[2]);
  v3 td = vscale(Iw, -(ka + ka * vnorm(w)));
  v3 f = vadd(fg, fd);
  return sv_make(vadd(td, vcross(l->com, f)), f);
}

/* Articulated-body algorithm (Featherstone RBDA table 7.1, floating base).  Outputs the
 * generalized acceleration in the [w_world, v_world, qdd] convention and fills the cache
 * used by impulse_response(). */
static int aba(const hrlo_env* E, const env_state* s, kin_t* K, const real tau[8], real udot[NDOF]) {
  const ant_model* M = &E->model;
  sv pA[NLINK], u_[NLINK];
  for (int i = 0; i < NLINK; i++) {
    sm I = spatial_inertia(&M->L[i]);
    K->IA[i] = I;
    sv Iv = sm_mulv(&I, K->v[i]);
    pA[i] = sv_sub(crf(K->v[i], Iv), link_ext_force(&E->cfg, &M->L[i], &K->Rw[i], K->v[i]));
  }
  for (int i = NLINK - 1; i >= 1; i--) {
    const link_t* l = &M->L[i];
    sm Ia = K->IA[i]; sv pa;
    if (l->jtype == 1) {
      sv S = sv_make(l->axis, V3(0, 0, 0));
      K->U[i] = sm_mulv(&K->IA[i], S);
      real D = sv_dot(S, K->U[i]);
      K->Dinv[i] = 1 / D;
      real ui = tau[l->dof] - sv_dot(S, pA[i]);
      u_[i].v[0] = ui;
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) Ia.m[a][b] -= K->U[i].v[a] * K->U[i].v[b] * K->Dinv[i];
      pa = sv_add(sv_add(pA[i], sm_mulv(&Ia, K->c[i])), sv_scale(K->U[i], ui * K->Dinv[i]));
    } else {
      pa = sv_add(pA[i], sm_mulv(&Ia, K->c[i]));
    }
    /* parent += X^T Ia X, X^T pa */
    sm T; /* Ia X */
    for (int a = 0; a < 6; a++)
      for (int b = 0; b < 6; b++) { real x = 0; for (int k = 0; k < 6; k++) x += Ia.m[a][k] * K->X[i].m[k][b]; T.m[a][b] = x; }
    int p = l->parent;
    for (int a = 0; a < 6; a++)
      for (int b = 0; b < 6; b++) { real x = 0; for (int k = 0; k < 6; k++) x += K->X[i].m[k][a] * T.m[k][b]; K->IA[p].m[a][b] += x; }
    pA[p] = sv_add(pA[p], sm_tmulv(&K->X[i], pa));
  }
  if (chol6(&K->IA[0], &K->L0)) return -1;
  sv a[NLINK];
  a[0] = chol6_solve(&K->L0, sv_scale(pA[0], -1));
  for (int i = 1; i < NLINK; i++) {
    const link_t* l = &M->L[i];
    sv ap = sv_add(sm_mulv(&K->X[i], a[l->parent]), K->c[i]);
    if (l->jtype == 1) {
      real qdd = (u_[i].v[0] - sv_dot(K->U[i], ap)) * K->Dinv[i];
      udot[6 + l->dof] = qdd;
      a[i] = sv_add(ap, sv_make(vscale(l->axis, qdd), V3(0, 0, 0)));
    } else a[i] = ap;
  }
  /* base: spatial -> classical acceleration, local -> world */
  v3 wl = sv_ang(K->v[0]), vl = sv_lin(K->v[0]);
  v3 wd = mmulv(&K->Rw[0], sv_ang(a[0]));
  v3 vd = mmulv(&K->Rw[0], vadd(sv_lin(a[0]), vcross(wl, vl)));
  for (int i = 0; i < 3; i++) { udot[i] = wd.v[i]; udot[3 + i] = vd.v[i]; }
  return 0;
}

/* dv = M^-1 f for a generalized impulse f (base part in world coordinates): the ABA impulse
 * response Bullet evaluates per constraint row (calcAccelerationDeltasMultiDof). */
static void impulse_response(const hrlo_env* E, const kin_t* K, const real f[NDOF], real dv[NDOF]) {
  const ant_model* M = &E->model;
  sv pA[NLINK], a[NLINK]; real u_[NLINK];
  for (int i = 0; i < NLINK; i++) pA[i] = sv_zero();
  for (int i = NLINK - 1; i >= 1; i--) {
    const link_t* l = &M->L[i];
    sv pa = pA[i];
    if (l->jtype == 1) {
      sv S = sv_make(l->axis, V3(0, 0, 0));
      u_[i] = f[6 + l->dof] - sv_dot(S, pA[i]);
      pa = sv_add(pa, sv_scale(K->U[i], u_[i] * K->Dinv[i]));
    }
    pA[l->parent] = sv_add(pA[l->parent], sm_tmulv(&K->X[i], pa));
  }
  v3 tl = mtmulv(&K->Rw[0], V3(f[0], f[1], f[2])), fl = mtmulv(&K->Rw[0], V3(f[3], f[4], f[5]));
  a[0] = chol6_solve(&K->L0, sv_sub(sv_make(tl, fl), pA[0]));
  for (int i = 1; i < NLINK; i++) {
    const link_t* l = &M->L[i];
    sv ap = sm_mulv(&K->X[i], a[l->parent]);
    if (l->jtype == 1) {
      real qdd = (u_[i] - sv_dot(K->U[i], ap)) * K->Dinv[i];
      dv[6 + l->dof] = qdd;
      a[i] = sv_add(ap, sv_make(vscale(l->axis, qdd), V3(0, 0, 0)));
    } else a[i] = ap;
  }
  v3 wd = mmulv(&K->Rw[0], sv_ang(a[0])), vd = mmulv(&K->Rw[0], sv_lin(a[0]));
  for (int i = 0; i < 3; i++) { dv[i] = wd.v[i]; dv[3 + i] = vd.v[i]; }
}

/* ======================================================================================
 * contacts: sphere vs ground slab top / 4 wall inner faces / maze box
 * ====================================================================================== */
typedef struct { int sphere; int link; v3 n, P; real dist; real mu; int item; } contact_t;
#define MAX_CONTACT_PER_GROUP 4
#define MAX_CONTACTS 16

/* Bullet btPlaneSpace1 */
static void plane_space(v3 n, v3* p, v3* q) {
  if (R_FABS(n.v[2]) > (real)0.7071067811865475244) {
    real a = n.v[1] * n.v[1] + n.v[2] * n.v[2];
    real k = 1 / R_SQRT(a);
    *p = V3(0, -n.v[2] * k, n.v[1] * k);
    *q = V3(a * k, -n.v[0] * p->v[2], n.v[0] * p->v[1]);
  } else {
    real a = n.v[0] * n.v[0] + n.v[1] * n.v[1];
    real k = 1 / R_SQRT(a);
    *p = V3(-n.v[1] * k, n.v[0] * k, 0);
    *q = V3(-n.v[2] * p->v[1], n.v[2] * p->v[0], a * k);
  }
}

static int sphere_vs_box(v3 c, real r, const float lo[3], const float hi[3], v3* n, real* dist) {
  v3 q, d; int inside = 1;
  for (int i = 0; i < 3; i++) {
    real x = c.v[i];
    if (x < (real)lo[i]) { x = (real)lo[i]; inside = 0; }
    if (x > (real)hi[i]) { x = (real)hi[i]; inside = 0; }
    q.v[i] = x;
  }
  if (!inside) {
    d = vsub(c, q);
    real len = vnorm(d);
    *n = vscale(d, 1 / len);
    *dist = len - r;
    return 1;
  }
  /* centre inside the box: exit through the nearest face */
  real best = (real)1e30; int bi = 0, bs = 1;
  for (int i = 0; i < 3; i++) {
    real dl = c.v[i] - (real)lo[i], dh = (real)hi[i] - c.v[i];
    if (dl < best) { best = dl; bi = i; bs = -1; }
    if (dh < best) { best = dh; bi = i; bs = +1; }
  }
  *n = V3(0, 0, 0); n->v[bi] = (real)bs;
  *dist = -best - r;
  return 1;
}

/* `touched` (optional, [HRL_MAX_ITEMS]): contact points per food/poison cube, what
 * getContactPoints(robot) reports after the step (ant_gather_env.py:114) */
/* The CYLINDER part of a capsule (segment A..B, radius r) against an axis-aligned box: the point of the segment's
 * INTERIOR that is closest to the box.  The squared distance from A + t d to the box is convex in t with the continuous,
 * monotone derivative g(t) = 2 sum_i e_i(t) d_i (e = the point's excess over the box, per axis); a minimum strictly
 * inside (0, 1) exists iff g(0) < 0 < g(1) and is bracketed by 24 bisection steps.  (A minimum at an end is the end
 * sphere's business.)  Normal and distance at that point are those of a sphere of radius r centred there. */
static real seg_box_grad(v3 A, v3 d, const float lo[3], const float hi[3], real t) {
  real g = 0;
  for (int i = 0; i < 3; i++) {
    real x = A.v[i] + t * d.v[i];
    real e = x > (real)hi[i] ? x - (real)hi[i] : (x < (real)lo[i] ? x - (real)lo[i] : 0);
    g += e * d.v[i];
  }
  return g;
}
static int capsule_interior_vs_box(v3 A, v3 B, real r, const float lo[3], const float hi[3], v3* Q, v3* n, real* dist) {
  v3 d = vsub(B, A);
  if (!(seg_box_grad(A, d, lo, hi, 0) < 0 && seg_box_grad(A, d, lo, hi, 1) > 0)) return 0;
  real a = 0, b = 1;
  for (int it = 0; it < 24; it++) {
    real m = (a + b) / 2;
    if (seg_box_grad(A, d, lo, hi, m) < 0) a = m; else b = m;
  }
  *Q = vadd(A, vscale(d, (a + b) / 2));
  return sphere_vs_box(*Q, r, lo, hi, n, dist);
}

static int detect_contacts(const hrlo_env* E, const env_state* s, const kin_t* K, contact_t* C, int feet_ground[4],
                           int* touched) {
  const hrl_config* cfg = &E->cfg;
  const ant_model* M = &E->model;
  int n = 0, per_group[4] = {0, 0, 0, 0};
  real margin = (real)cfg->contact_margin;
  real wx = (real)cfg->world_size[0] / 2 - (real)0.05, wy = (real)cfg->world_size[1] / 2 - (real)0.05;
  for (int k = 0; k < 4; k++) feet_ground[k] = 0;
  for (int si = 0; si < NSPHERE; si++) {
    const sphere_t* S = &M->S[si];
    v3 c = vadd(K->ow[S->link], mmulv(&K->Rw[S->link], S->local));
    real r = S->radius;
    for (int surf = 0; surf < 6; surf++) {
      v3 nrm; real dist;
      switch (surf) {
        case 0: nrm = V3(0, 0, 1); dist = c.v[2] - (real)cfg->ground_z - r; break;
        case 1: if (!cfg->has_walls) continue; nrm = V3(-1, 0, 0); dist = wx - c.v[0] - r; break;
        case 2: if (!cfg->has_walls) continue; nrm = V3(1, 0, 0); dist = c.v[0] + wx - r; break;
        case 3: if (!cfg->has_walls) continue; nrm = V3(0, -1, 0); dist = wy - c.v[1] - r; break;
        case 4: if (!cfg->has_walls) continue; nrm = V3(0, 1, 0); dist = c.v[1] + wy - r; break;
        default: if (!cfg->has_box) continue; sphere_vs_box(c, r, cfg->box_lo, cfg->box_hi, &nrm, &dist); break;
      }
      if (!(dist < margin)) continue;
      if (surf == 0 && S->foot >= 0) feet_ground[S->foot] = 1;
      if (per_group[S->group] >= MAX_CONTACT_PER_GROUP) continue;
      per_group[S->group]++;
      C[n].sphere = si; C[n].link = S->link; C[n].n = nrm; C[n].dist = dist; C[n].P = vsub(c, vscale(nrm, r));
      C[n].mu = (real)cfg->friction; C[n].item = -1;
      n++;
    }
    if (cfg->item_contacts) { /* food / poison cubes: static axis-aligned boxes (assets/food.xml, gather_scene.py:62) */
      real hh = (real)cfg->item_half, reach = hh + r + margin;
      for (int i = 0; i < cfg->n_food + cfg->n_poison; i++) {
        real dx = c.v[0] - s->items[i][0], dy = c.v[1] - s->items[i][1];
        if (R_FABS(dx) > reach || R_FABS(dy) > reach) continue;
        float lo[3] = {(float)s->items[i][0] - cfg->item_half, (float)s->items[i][1] - cfg->item_half, cfg->item_z - cfg->item_half};
        float hi[3] = {(float)s->items[i][0] + cfg->item_half, (float)s->items[i][1] + cfg->item_half, cfg->item_z + cfg->item_half};
        v3 nrm; real dist;
        sphere_vs_box(c, r, lo, hi, &nrm, &dist);
        if (!(dist < margin)) continue;
        if (touched) touched[i]++;
        if (per_group[S->group] >= MAX_CONTACT_PER_GROUP) continue;
        per_group[S->group]++;
        C[n].sphere = si; C[n].link = S->link; C[n].n = nrm; C[n].dist = dist; C[n].P = vsub(c, vscale(nrm, r));
        C[n].mu = (real)cfg->item_friction; C[n].item = i;
        n++;
      }
    }
    /* Maze box: the CYLINDER part of the leg's three capsules against the box's four vertical edges (the ant walks
     * around the corners of the U, maze_scene.py:13; the end-spheres above only cover contacts at a capsule end).
     * For a segment outside a convex rectangle the closest pair is either (segment end, rectangle) - the spheres - or
     * (rectangle corner, segment interior) - this test; the legs never reach the box's top (z = 2), so it is planar.
     * Runs after the last sphere of leg k (hip), order: foot, aux, leg capsule x corners (lo,lo) (hi,lo) (lo,hi) (hi,hi). */
    if (cfg->has_box && si >= 1 && (si - 1) % 3 == 2) {
      int k = (si - 1) / 3;
      real sx = (real)LEG_SX[k], sy = (real)LEG_SY[k], rc = (real)ANT_R_CAPS;
      for (int cap = 0; cap < 3; cap++) {
        int link = 3 + 3 * k - cap;  /* foot, aux, leg */
        real len = (cap == 0) ? (real)0.4 : (real)0.2;
        v3 A = K->ow[link], B = vadd(A, mmulv(&K->Rw[link], V3(len * sx, len * sy, 0)));
        v3 d = vsub(B, A);
        real L2 = d.v[0] * d.v[0] + d.v[1] * d.v[1];
        if (!(L2 > (real)1e-12)) continue;
        for (int corner = 0; corner < 4; corner++) {
          real cx = (corner & 1) ? (real)cfg->box_hi[0] : (real)cfg->box_lo[0], sgx = (corner & 1) ? (real)1 : (real)-1;
          real cy = (corner & 2) ? (real)cfg->box_hi[1] : (real)cfg->box_lo[1], sgy = (corner & 2) ? (real)1 : (real)-1;
          real t = ((cx - A.v[0]) * d.v[0] + (cy - A.v[1]) * d.v[1]) / L2;
          if (!(t > 0 && t < 1)) continue;
          v3 Q = vadd(A, vscale(d, t));
          if (Q.v[2] < (real)cfg->box_lo[2] || Q.v[2] > (real)cfg->box_hi[2]) continue;
          real ex = Q.v[0] - cx, ey = Q.v[1] - cy;
          if (ex * sgx < 0 || ey * sgy < 0) continue;  /* not in the corner's Voronoi region: a face is closer */
          real el = R_SQRT(ex * ex + ey * ey);
          if (!(el > 0)) continue;
          real dist = el - rc;
          if (!(dist < margin)) continue;
          if (per_group[k] >= MAX_CONTACT_PER_GROUP) continue;
          per_group[k]++;
          v3 nrm = V3(ex / el, ey / el, 0);
          C[n].sphere = -1; C[n].link = link; C[n].n = nrm; C[n].dist = dist; C[n].P = vsub(Q, vscale(nrm, rc));
          C[n].mu = (real)cfg->friction; C[n].item = -1;
          n++;
        }
      }
    }
    /* Food / poison cubes: the cylinder part of the leg's three capsules (assets/ant.xml:16-24 capsules vs the 0.25 m
     * boxes of assets/food.xml:17-22) - a leg lying across a cube's edge touches it between its end spheres.  After the
     * last sphere of leg k; cubes in index order, capsules foot, aux, leg.  Each such contact is a contact POINT of the
     * robot with the cube (ant_gather_env.py:113-116 counts them). */
    if (cfg->item_contacts && si >= 1 && (si - 1) % 3 == 2) {
      int k = (si - 1) / 3;
      real sx = (real)LEG_SX[k], sy = (real)LEG_SY[k], rc = (real)ANT_R_CAPS, hh = (real)cfg->item_half, rr = hh + rc + margin;
      for (int i = 0; i < cfg->n_food + cfg->n_poison; i++) {
        real bx = s->items[i][0], by = s->items[i][1];
        float lo[3] = {(float)bx - cfg->item_half, (float)by - cfg->item_half, cfg->item_z - cfg->item_half};
        float hi[3] = {(float)bx + cfg->item_half, (float)by + cfg->item_half, cfg->item_z + cfg->item_half};
        for (int cap = 0; cap < 3; cap++) {
          int link = 3 + 3 * k - cap;  /* foot, aux, leg */
          real len = (cap == 0) ? (real)0.4 : (real)0.2;
          v3 A = K->ow[link], B = vadd(A, mmulv(&K->Rw[link], V3(len * sx, len * sy, 0)));
          real x0 = A.v[0] < B.v[0] ? A.v[0] : B.v[0], x1 = A.v[0] < B.v[0] ? B.v[0] : A.v[0];
          real y0 = A.v[1] < B.v[1] ? A.v[1] : B.v[1], y1 = A.v[1] < B.v[1] ? B.v[1] : A.v[1];
          if (bx < x0 - rr || bx > x1 + rr || by < y0 - rr || by > y1 + rr) continue;
          v3 Q, nrm; real dist;
          if (!capsule_interior_vs_box(A, B, rc, lo, hi, &Q, &nrm, &dist)) continue;
          if (!(dist < margin)) continue;
          if (touched) touched[i]++;
          if (per_group[k] >= MAX_CONTACT_PER_GROUP) continue;
          per_group[k]++;
          C[n].sphere = -1; C[n].link = link; C[n].n = nrm; C[n].dist = dist; C[n].P = vsub(Q, vscale(nrm, rc));
          C[n].mu = (real)cfg->item_friction; C[n].item = i;
          n++;
        }
      }
    }
  }
  return n;
}

/* Jacobian row of direction d at world point P on `link`: J u = d . v_P */
static void point_jacobian(const ant_model* M, const kin_t* K, int link, v3 P, v3 d, real J[NDOF]) {
  for (int i = 0; i < NDOF; i++) J[i] = 0;
  v3 t = vcross(vsub(P, K->ow[0]), d);
  for (int i = 0; i < 3; i++) { J[i] = t.v[i]; J[3 + i] = d.v[i]; }
  for (int i = link; i > 0; i = M->L[i].parent)
    if (M->L[i].jtype == 1) {
      v3 a = mmulv(&K->Rw[i], M->L[i].axis);
      J[6 + M->L[i].dof] = vdot(a, vcross(vsub(P, K->ow[i]), d));
    }
}

/* ======================================================================================
 * constraint rows + projected Gauss-Seidel
 * ====================================================================================== */
typedef struct { real J[NDOF], W[NDOF], diagInv, rhs, lo, hi, lam; } row_t;

static void row_finish(const hrlo_env* E, const kin_t* K, row_t* r, const real u[NDOF], real pen, real erp, real h,
                       int positional) {
  impulse_response(E, K, r->J, r->W);
  real d = 0, rel = 0;
  for (int i = 0; i < NDOF; i++) { d += r->J[i] * r->W[i]; rel += r->J[i] * u[i]; }
  r->diagInv = 1 / d;
  real posErr = 0, velErr = -rel;
  if (positional) {
    if (pen > 0) velErr -= pen / h;
    else posErr = -pen * erp / h;
  }
  r->rhs = (posErr + velErr) * r->diagInv;
  r->lam = 0;
}

static void row_apply(row_t* r, real dv[NDOF], real dl) {
  r->lam += dl;
  for (int i = 0; i < NDOF; i++) dv[i] += r->W[i] * dl;
}
static real row_delta(const row_t* r, const real dv[NDOF]) {
  real jd = 0;
  for (int i = 0; i < NDOF; i++) jd += r->J[i] * dv[i];
  return r->rhs - jd * r->diagInv;
}

static void clamp_vel(real u[NDOF], real vmax) {
  for (int i = 0; i < NDOF; i++) { if (u[i] > vmax) u[i] = vmax; if (u[i] < -vmax) u[i] = -vmax; }
}

static void integrate_pose(env_state* s, const real u[NDOF], real h) {
  for (int i = 0; i < 3; i++) { s->ang[i] = u[i]; s->vel[i] = u[3 + i]; s->pos[i] += h * u[3 + i]; }
  for (int j = 0; j < 8; j++) { s->qd[j] = u[6 + j]; s->q[j] += h * u[6 + j]; }
  /* q <- exp(w h) * q   (Bullet pQuatUpdateFun, world-frame omega) */
  real w[3] = {u[0], u[1], u[2]};
  real ang = R_SQRT(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  real ax[3], cw;
  if (ang * h > (real)(0.25 * PI_D)) ang = (real)(0.25 * PI_D) / h; /* ANGULAR_MOTION_THRESHOLD */
  real k;
  if (ang < (real)0.001) k = (real)0.5 * h - h * h * h * (real)0.020833333333 * ang * ang;
  else k = R_SIN((real)0.5 * ang * h) / ang;
  for (int i = 0; i < 3; i++) ax[i] = w[i] * k;
  cw = R_COS((real)0.5 * ang * h);
  real x = s->quat[0], y = s->quat[1], z = s->quat[2], qw = s->quat[3];
  real nx = cw * x + ax[0] * qw + ax[1] * z - ax[2] * y;
  real ny = cw * y + ax[1] * qw + ax[2] * x - ax[0] * z;
  real nz = cw * z + ax[2] * qw + ax[0] * y - ax[1] * x;
  real nw = cw * qw - ax[0] * x - ax[1] * y - ax[2] * z;
  real inv = 1 / R_SQRT(nx * nx + ny * ny + nz * nz + nw * nw);
  s->quat[0] = nx * inv; s->quat[1] = ny * inv; s->quat[2] = nz * inv; s->quat[3] = nw * inv;
}

/* one Bullet internal step of h = dt/substeps for the ant.  feet_ground = foot-vs-floor
 * manifold exists at the START of this sub-step (collision detection precedes dynamics). */
static int ant_substep(hrlo_env* E, env_state* s, const real tau[8], int feet_ground[4], int* touched) {
  const hrl_config* cfg = &E->cfg;
  const ant_model* M = &E->model;
  real h = (real)cfg->dt / (real)cfg->substeps;
  kin_t K;
  forward_kinematics(M, s, &K);
  contact_t C[MAX_CONTACTS];
  int nc = detect_contacts(E, s, &K, C, feet_ground, touched);
  real u[NDOF], udot[NDOF];
  for (int i = 0; i < 3; i++) { u[i] = s->ang[i]; u[3 + i] = s->vel[i]; }
  for (int j = 0; j < 8; j++) u[6 + j] = s->qd[j];
  if (aba(E, s, &K, tau, udot)) return -1;
  for (int i = 0; i < NDOF; i++) u[i] += h * udot[i];
  clamp_vel(u, (real)cfg->max_coord_vel);

  /* rows: joint limits (only when violated), contact normals, friction pairs */
  row_t lim[8], nrm[MAX_CONTACTS], fr[2 * MAX_CONTACTS];
  int nl = 0;
  for (int j = 0; j < 8; j++) {
    real pen_lo = s->q[j] - M->lo[j], pen_hi = M->hi[j] - s->q[j];
    for (int side = 0; side < 2; side++) {
      real pen = side ? pen_hi : pen_lo;
      if (pen > 0) continue;
      row_t* r = &lim[nl++];
      memset(r->J, 0, sizeof r->J);
      r->J[6 + j] = side ? (real)-1 : (real)1;
      r->lo = 0; r->hi = (real)cfg->limit_max_impulse;
      row_finish(E, &K, r, u, pen, (real)cfg->limit_erp, h, 1);
    }
  }
  for (int c = 0; c < nc; c++) {
    int link = C[c].link;
    row_t* r = &nrm[c];
    point_jacobian(M, &K, link, C[c].P, C[c].n, r->J);
    r->lo = 0; r->hi = (real)1e10;
    row_finish(E, &K, r, u, C[c].dist, (real)cfg->contact_erp, h, 1);
    v3 t1, t2; plane_space(C[c].n, &t1, &t2);
    point_jacobian(M, &K, link, C[c].P, t1, fr[2 * c].J);
    point_jacobian(M, &K, link, C[c].P, t2, fr[2 * c + 1].J);
    row_finish(E, &K, &fr[2 * c], u, 0, 0, h, 0);
    row_finish(E, &K, &fr[2 * c + 1], u, 0, 0, h, 0);
  }
  E->n_contacts += nc; E->n_limit_rows += nl; E->n_substeps += 1;
  for (int c = 0; c < nc; c++) if (C[c].sphere < 0) E->n_capsule_contacts += 1;

  real dv[NDOF];
  for (int i = 0; i < NDOF; i++) dv[i] = 0;
  for (int it = 0; it < cfg->solver_iters; it++) {
    for (int jj = 0; jj < nl; jj++) {
      int j = (it & 1) ? jj : nl - 1 - jj; /* Bullet alternates the non-contact row order */
      row_t* r = &lim[j];
      real dl = row_delta(r, dv), sum = r->lam + dl;
      if (sum < r->lo) dl = r->lo - r->lam; else if (sum > r->hi) dl = r->hi - r->lam;
      row_apply(r, dv, dl);
    }
    for (int c = 0; c < nc; c++) {
      row_t* r = &nrm[c];
      real dl = row_delta(r, dv), sum = r->lam + dl;
      if (sum < r->lo) dl = r->lo - r->lam; else if (sum > r->hi) dl = r->hi - r->lam;
      row_apply(r, dv, dl);
    }
    for (int c = 0; c < nc; c++) { /* implicit-cone friction pair */
      real tot = nrm[c].lam;
      if (!(tot > 0)) continue;
      row_t *a = &fr[2 * c], *b = &fr[2 * c + 1];
      real da = row_delta(a, dv), db = row_delta(b, dv);
      real sa = a->lam + da, sb = b->lam + db, lim2 = C[c].mu * tot;
      real len2 = sa * sa + sb * sb;
      if (len2 > lim2 * lim2) { real sc = lim2 / R_SQRT(len2); sa *= sc; sb *= sc; }
      da = sa - a->lam; db = sb - b->lam;
      row_apply(a, dv, da); row_apply(b, dv, db);
    }
  }
  for (int i = 0; i < NDOF; i++) u[i] += dv[i];
  clamp_vel(u, (real)cfg->max_coord_vel);
  integrate_pose(s, u, h);
  return 0;
}

/* ======================================================================================
 * PointGather body (point_bot.py, assets/player_cube.xml): the north star's "point-mass
 * integrator": a 10 kg, 0.35 half-extent cube that only translates.
 * ====================================================================================== */
#define POINT_MASS 10.0  /* player_cube.xml:8 */
#define POINT_HALF 0.35  /* player_cube.xml:8 */
static void point_substep(hrlo_env* E, env_state* s, const real force[3]) {
  const hrl_config* cfg = &E->cfg;
  real h = (real)cfg->dt / (real)cfg->substeps, m = (real)POINT_MASS, half = (real)POINT_HALF;
  real wx = (real)cfg->world_size[0] / 2 - (real)0.05 - half, wy = (real)cfg->world_size[1] / 2 - (real)0.05 - half;
  /* contacts from start-of-substep position */
  v3 N[5]; real D[5]; int nc = 0;
  real margin = (real)cfg->contact_margin;
  real dg = s->pos[2] - half - (real)cfg->ground_z;
  if (dg < margin) { N[nc] = V3(0, 0, 1); D[nc++] = dg; }
  if (cfg->has_walls) {
    if (wx - s->pos[0] < margin) { N[nc] = V3(-1, 0, 0); D[nc++] = wx - s->pos[0]; }
    if (s->pos[0] + wx < margin) { N[nc] = V3(1, 0, 0); D[nc++] = s->pos[0] + wx; }
    if (wy - s->pos[1] < margin) { N[nc] = V3(0, -1, 0); D[nc++] = wy - s->pos[1]; }
    if (s->pos[1] + wy < margin) { N[nc] = V3(0, 1, 0); D[nc++] = s->pos[1] + wy; }
  }
  v3 v = V3(s->vel[0], s->vel[1], s->vel[2]);
  real kl = (real)cfg->lin_damping;
  v3 f = V3(force[0], force[1], force[2] - m * (real)cfg->gravity);
  f = vadd(f, vscale(v, -m * (kl + kl * vnorm(v))));
  v = vadd(v, vscale(f, h / m));
  for (int i = 0; i < 3; i++) { if (v.v[i] > (real)cfg->max_coord_vel) v.v[i] = (real)cfg->max_coord_vel; if (v.v[i] < -(real)cfg->max_coord_vel) v.v[i] = -(real)cfg->max_coord_vel; }
  real lamn[5] = {0}, lama[5] = {0}, lamb[5] = {0}, rhsn[5], rhsa[5], rhsb[5];
  v3 T1[5], T2[5];
  for (int c = 0; c < nc; c++) {
    real rel = vdot(N[c], v), posErr = 0, velErr = -rel;
    if (D[c] > 0) velErr -= D[c] / h; else posErr = -D[c] * (real)cfg->contact_erp / h;
    rhsn[c] = (posErr + velErr) * m;
    plane_space(N[c], &T1[c], &T2[c]);
    rhsa[c] = -vdot(T1[c], v) * m; rhsb[c] = -vdot(T2[c], v) * m;
  }
  v3 dv = V3(0, 0, 0);
  real mu = (real)cfg->friction;
  for (int it = 0; it < cfg->solver_iters; it++) {
    for (int c = 0; c < nc; c++) {
      real dl = rhsn[c] - vdot(N[c], dv) * m, sum = lamn[c] + dl;
      if (sum < 0) dl = -lamn[c];
      lamn[c] += dl; dv = vadd(dv, vscale(N[c], dl / m));
    }
    for (int c = 0; c < nc; c++) {
      if (!(lamn[c] > 0)) continue;
      real da = rhsa[c] - vdot(T1[c], dv) * m, db = rhsb[c] - vdot(T2[c], dv) * m;
      real sa = lama[c] + da, sb = lamb[c] + db, lim = mu * lamn[c], len2 = sa * sa + sb * sb;
      if (len2 > lim * lim) { real sc = lim / R_SQRT(len2); sa *= sc; sb *= sc; }
      da = sa - lama[c]; db = sb - lamb[c];
      lama[c] = sa; lamb[c] = sb;
      dv = vadd(dv, vadd(vscale(T1[c], da / m), vscale(T2[c], db / m)));
    }
  }
  v = vadd(v, dv);
  for (int i = 0; i < 3; i++) { s->vel[i] = v.v[i]; s->pos[i] += h * v.v[i]; }
}

/* ======================================================================================
 * task layer
 * ====================================================================================== */
/* The config carries spans as float32; the reference defaults are the float64 constants pi and
 * 2*pi (ant_gather_env.py:22, ant_maze_bullet_env.py:23): snap those two back. */
static double snap_span(float s) {
  if (s == (float)(2.0 * PI_D)) return 2.0 * PI_D;
  if (s == (float)PI_D) return PI_D;
  return (double)s;
}

/* Bullet getEulerFromQuaternion (SURVEY.md A.3 "Queries") */
static void quat_to_rpy(const real q[4], real rpy[3]) {
  real x = q[0], y = q[1], z = q[2], w = q[3];
  real sarg = -2 * (x * z - w * y);
  if (sarg <= (real)-0.99999) { rpy[1] = (real)(-0.5 * PI_D); rpy[0] = 0; rpy[2] = 2 * R_ATAN2(x, -y); }
  else if (sarg >= (real)0.99999) { rpy[1] = (real)(0.5 * PI_D); rpy[0] = 0; rpy[2] = 2 * R_ATAN2(-x, y); }
  else {
    rpy[1] = R_ASIN(sarg);
    rpy[0] = R_ATAN2(2 * (y * z + w * x), w * w - x * x - y * y + z * z);
    rpy[2] = R_ATAN2(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
  }
}

/* Gather sector sensor - ant_gather_env.py:128-177 (twin gather_base.py:118-168), always in
 * double like the reference.  bins_out[i] = bin of item i or -1. */
void hrlo_gather_sensor_one(int n_bins, double sensor_range, double span, double rx, double ry, double yaw,
                            const double* items_xy, int n_food, int n_poison, const double* d2_in, double* food,
                            double* poison, int32_t* bins_out) {
  int n = n_food + n_poison, order[HRL_MAX_ITEMS];
  double d2[HRL_MAX_ITEMS];
  for (int i = 0; i < n_bins; i++) food[i] = poison[i] = 0.0;
  for (int i = 0; i < n; i++) {
    double dx = items_xy[2 * i] - rx, dy = items_xy[2 * i + 1] - ry;
    d2[i] = d2_in ? d2_in[i] : dx * dx + dy * dy; /* ant_gather_env.py:198-200 (squared!) */
    order[i] = i;
    if (bins_out) bins_out[i] = -1;
  }
  /* sorted(..., key=dist, reverse=True): stable, descending (ant_gather_env.py:141) */
  for (int i = 1; i < n; i++) {
    int o = order[i], j = i - 1;
    while (j >= 0 && d2[order[j]] < d2[o]) { order[j + 1] = order[j]; j--; }
    order[j + 1] = o;
  }
  double bin_res = span / n_bins, half_span = span * 0.5;
  for (int k = 0; k < n; k++) {
    int i = order[k];
    if (d2[i] > sensor_range) continue; /* :145 */
    double angle = atan2(items_xy[2 * i + 1] - ry, items_xy[2 * i] - rx) - yaw; /* :148 */
    angle = fmod(angle, 2 * PI_D);
    if (angle < 0) angle += 2 * PI_D; /* Python % semantics (:151) */
    if (angle > PI_D) angle -= 2 * PI_D;
    if (angle < -PI_D) angle += 2 * PI_D;
    if (fabs(angle) > half_span) continue; /* :159 */
    int b = (int)((angle + half_span) / bin_res); /* :161 */
    if (b > n_bins - 1) b = n_bins - 1; /* reference raises IndexError at exactly +half_span; clamp */
    double inten = 1.0 - d2[i] / sensor_range; /* :162 */
    if (i < n_food) food[b] = inten; else poison[b] = inten;
    if (bins_out) bins_out[i] = b;
  }
}

/* intersection_utils.py:93-104 */
static int quadrant_d(double x, double y) {
  if (x >= 0 && y >= 0) return 1;
  if (x >= 0 && y <= 0) return 4;
  if (x <= 0 && y >= 0) return 2;
  if (x <= 0 && y <= 0) return 3;
  return 0;
}
int hrlo_quadrant(double x, double y) { return quadrant_d(x, y); }

/* intersection_utils.py:84-90 */
int hrlo_find_intersection(const double p[8], double out[2]) {
  double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3], x3 = p[4], y3 = p[5], x4 = p[6], y4 = p[7];
  double d = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
  if (d == 0) return 0;
  out[0] = ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4)) / d;
  out[1] = ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4)) / d;
  return 1;
}

/* intersection_utils.py:14-71 */
static int orient_d(const double* p, const double* q, const double* r) {
  double val = ((q[1] - p[1]) * (r[0] - q[0])) - ((q[0] - p[0]) * (r[1] - q[1]));
  return val > 0 ? 1 : (val < 0 ? 2 : 0);
}
static int on_seg_d(const double* p, const double* q, const double* r) {
  return q[0] <= fmax(p[0], r[0]) && q[0] >= fmin(p[0], r[0]) && q[1] <= fmax(p[1], r[1]) && q[1] >= fmin(p[1], r[1]);
}
int hrlo_segment_intersection(const double s[8]) {
  const double *p1 = s, *q1 = s + 2, *p2 = s + 4, *q2 = s + 6;
  int o1 = orient_d(p1, q1, p2), o2 = orient_d(p1, q1, q2), o3 = orient_d(p2, q2, p1), o4 = orient_d(p2, q2, q1);
  if (o1 != o2 && o3 != o4) return 1;
  if (o1 == 0 && on_seg_d(p1, p2, q1)) return 1;
  if (o2 == 0 && on_seg_d(p1, q2, q1)) return 1;
  if (o3 == 0 && on_seg_d(p2, p1, q2)) return 1;
  if (o4 == 0 && on_seg_d(p2, q1, q2)) return 1;
  return 0;
}

/* Wall lidar - sizeable_enclosed_scene.py:63-97, in double like the reference. */
void hrlo_sense_walls_one(int n_bins, double span, double range, int n_lines, const double* bounds, double px,
                          double py, double yaw, double* out) {
  for (int i = 0; i < n_bins; i++) {
    double ang;
    if (span == 2 * PI_D) ang = PI_D / 2 + yaw + ((double)(i + 1) / n_bins) * span; /* :68-69 */
    else ang = PI_D / 2 + yaw + ((double)i / (n_bins - 1)) * span;                 /* :70-71 */
    double sx = px + range * cos(ang), sy = py + range * sin(ang); /* :72 + pol2cart */
    int sq = quadrant_d(sx - px, sy - py);                         /* :74 */
    double best = 0;
    for (int l = 0; l < n_lines; l++) {
      double p[8] = {px, py, sx, sy, bounds[4 * l], bounds[4 * l + 1], bounds[4 * l + 2], bounds[4 * l + 3]}, in[2];
      if (!hrlo_find_intersection(p, in)) continue;                /* :81-83 */
      double dx = px - in[0], dy = py - in[1];
      double dist = sqrt(dx * dx + dy * dy);                       /* :85 */
      if (dist > range) continue;                                  /* :86 */
      if (sq != quadrant_d(in[0] - px, in[1] - py)) continue;      /* :89 */
      double val = 1. - dist / range;
      if (val > best) best = val;                                  /* :95 */
    }
    out[i] = best;
  }
}

/* Maze goal sector sensor - ant_maze_bullet_env.py:135-178, in double like the reference.
 * `wtd` is the (Q1-distorted) robot.walk_target_dist, the pose is the true torso pose; the goal is
 * hidden when the segment robot->goal crosses one of the box_bounds segments (maze_scene.py:19-21). */
void hrlo_maze_target_sensor(int n_bins, double span, double range, int n_box, const double* box_bounds, double rx,
                             double ry, double yaw, double tx, double ty, double wtd, double* readings) {
  for (int b = 0; b < n_bins; b++) readings[b] = 0;
  if (n_bins <= 0) return;
  if (wtd > range) return; /* :145 */
  for (int l = 0; l < n_box; l++) {
    const double sgm[8] = {rx, ry, tx, ty, box_bounds[4 * l], box_bounds[4 * l + 1], box_bounds[4 * l + 2], box_bounds[4 * l + 3]};
    if (hrlo_segment_intersection(sgm)) return; /* :148-150 */
  }
  double ang = atan2(ty - ry, tx - rx) - yaw; /* :153 */
  ang = fmod(ang, 2 * PI_D);
  if (ang < 0) ang += 2 * PI_D; /* Python % */
  if (ang > PI_D) ang -= 2 * PI_D;
  if (ang < -PI_D) ang += 2 * PI_D;
  double half = span * 0.5, res = span / n_bins;
  if (fabs(ang) > half) return; /* :164 */
  int bin = (int)((ang + half) / res);
  if (bin > n_bins - 1) bin = n_bins - 1; /* the reference raises IndexError at exactly +half span */
  readings[bin] = 1.0 - wtd / range; /* :167 */
}

/* gather_scene.py:52-62; returns number of (x,y) attempts consumed.  Attempt k of this
 * placement is Philox(draw, env, stream, item*64+k); replay != NULL replays uniforms instead. */
static int random_on_plane(const hrl_config* cfg, real ax, real ay, uint64_t seed, uint32_t env, uint32_t stream,
                           uint32_t draw, int item, const double* replay, real out[2]) {
  real sx = (real)cfg->world_size[0] - 1, sy = (real)cfg->world_size[1] - 1;
  int used = 0;
  for (;;) {
    real u[4];
    if (replay) { u[0] = (real)replay[2 * used]; u[1] = (real)replay[2 * used + 1]; }
    else rng_u4(seed, env, stream, draw, (uint32_t)(item * 64 + used), u);
    used++;
    real x = u[0] * sx - sx / 2, y = u[1] * sy - sy / 2;
    real dx = ax - x, dy = ay - y;
    int last = replay ? (used >= 1000) : (used >= MAX_PLACE_ATTEMPTS);
    if (R_SQRT(dx * dx + dy * dy) < (real)cfg->robot_object_spacing && !last) continue;
    out[0] = x; out[1] = y;
    return used;
  }
}
/* test hook: the rejection rule with replayed uniforms (tests/golden/random_on_plane.npz) */
int hrlo_random_on_plane_replay(double world_x, double world_y, double spacing, double ax, double ay,
                                const double* uniforms, double* out_xy) {
  hrl_config c; memset(&c, 0, sizeof c);
  c.world_size[0] = (float)world_x; c.world_size[1] = (float)world_y; c.robot_object_spacing = (float)spacing;
  real o[2];
  int used = random_on_plane(&c, (real)ax, (real)ay, 0, 0, 0, 0, 0, uniforms, o);
  out_xy[0] = (double)o[0]; out_xy[1] = (double)o[1];
  return 2 * used;
}

typedef struct {
  real obs28[28];
  real rpy[3];
  real body_xy[2];
  int joints_at_limit;
  real relpos[8], jspeed[8];
  real wtd; /* walk_target_dist of THIS calc_state */
} calc_t;

static real clip5(real x) { return x > 5 ? 5 : (x < -5 ? -5 : x); }

/* WalkerBase.calc_state [3P-MEM]: SURVEY.md App. A.2 / B.1 */
static void ant_calc_state(const hrlo_env* E, const env_state* s, calc_t* c) {
  const ant_model* M = &E->model;
  const hrl_config* cfg = &E->cfg;
  kin_t K;
  forward_kinematics(M, s, &K);
  c->joints_at_limit = 0;
  for (int j = 0; j < 8; j++) {
    real mid = (real)0.5 * (M->lo[j] + M->hi[j]);
    c->relpos[j] = 2 * (s->q[j] - mid) / (M->hi[j] - M->lo[j]);
    c->jspeed[j] = (real)0.1 * s->qd[j];
    if (R_FABS(c->relpos[j]) > (real)0.99) c->joints_at_limit++;
  }
  real sx = 0, sy = 0;
  for (int i = 0; i < NLINK; i++) { sx += K.comw[i].v[0]; sy += K.comw[i].v[1]; }
  /* quirk Q1: scene bodies are averaged in as well */
  sx += (real)cfg->scene_parts_sum[0]; sy += (real)cfg->scene_parts_sum[1];
  real np = (real)(NLINK + cfg->n_scene_parts);
  c->body_xy[0] = sx / np; c->body_xy[1] = sy / np;
  quat_to_rpy(s->quat, c->rpy);
  real z = s->pos[2];
  real yaw = c->rpy[2];
  real theta = R_ATAN2(s->target[1] - c->body_xy[1], s->target[0] - c->body_xy[0]);
  real dx = s->target[0] - c->body_xy[0], dy = s->target[1] - c->body_xy[1];
  c->wtd = R_SQRT(dx * dx + dy * dy);
  real a2t = theta - yaw;
  real cy = R_COS(-yaw), sy_ = R_SIN(-yaw);
  real vx = cy * s->vel[0] - sy_ * s->vel[1], vy = sy_ * s->vel[0] + cy * s->vel[1], vz = s->vel[2];
  real* o = c->obs28;
  o[0] = z - s->initial_z; o[1] = R_SIN(a2t); o[2] = R_COS(a2t);
  o[3] = (real)0.3 * vx; o[4] = (real)0.3 * vy; o[5] = (real)0.3 * vz; o[6] = c->rpy[0]; o[7] = c->rpy[1];
  for (int j = 0; j < 8; j++) { o[8 + 2 * j] = c->relpos[j]; o[9 + 2 * j] = c->jspeed[j]; }
  for (int k = 0; k < 4; k++) o[24 + k] = s->feet[k];
  for (int i = 0; i < 28; i++) o[i] = clip5(o[i]);
}

static int all_finite(const real* x, int n) {
  for (int i = 0; i < n; i++) if (!isfinite((double)x[i])) return 0;
  return 1;
}

static void items_to_double(const env_state* s, double* it) {
  for (int i = 0; i < HRL_MAX_ITEMS; i++) { it[2 * i] = (double)s->items[i][0]; it[2 * i + 1] = (double)s->items[i][1]; }
}

static int imin(int a, int b) { return a < b ? a : b; }
/* width of the food/poison part of a Gather observation: 2 n_bins sector readings, or with
 * use_sensor=False the xy of the min(n_bins, n) nearest food and poison items
 * (ant_gather_env.py:179-196; the reference's observation_space keeps the sensor width, :54-55) */
static int food_obs_dim(const hrl_config* c) {
  if (c->use_sensor) return 2 * c->n_bins;
  return 2 * imin(c->n_bins, c->n_food) + 2 * imin(c->n_bins, c->n_poison);
}
/* get_food_obs (ant_gather_env.py:120-124): sector sensor or get_abs_pos (:179-196, twin
 * gather_base.py:170-187): items of each type sorted by squared distance (stable), first n_bins, world xy */
static void food_obs(const hrl_config* cfg, const env_state* s, real yaw, real* out) {
  if (cfg->use_sensor) {
    double it[2 * HRL_MAX_ITEMS], fo[HRL_MAX_BINS], po[HRL_MAX_BINS];
    items_to_double(s, it);
    /* items are stored food-first; with n_food < 8 the poison block still starts at index n_food */
    hrlo_gather_sensor_one(cfg->n_bins, cfg->sensor_range, snap_span(cfg->sensor_span), (double)s->pos[0], (double)s->pos[1],
                           (double)yaw, it, cfg->n_food, cfg->n_poison, NULL, fo, po, NULL);
    for (int b = 0; b < cfg->n_bins; b++) { out[b] = (real)fo[b]; out[cfg->n_bins + b] = (real)po[b]; }
    return;
  }
  int o = 0;
  for (int ty = 0; ty < 2; ty++) {
    int first = ty ? cfg->n_food : 0, n = ty ? cfg->n_poison : cfg->n_food, keep = imin(cfg->n_bins, n);
    double d2[8]; int idx[8];
    for (int i = 0; i < n; i++) {
      double dx = (double)s->items[first + i][0] - (double)s->pos[0], dy = (double)s->items[first + i][1] - (double)s->pos[1];
      d2[i] = dx * dx + dy * dy; idx[i] = i;
    }
    for (int i = 1; i < n; i++) { /* stable insertion sort = Python sorted() */
      int k = idx[i], j = i - 1;
      while (j >= 0 && d2[idx[j]] > d2[k]) { idx[j + 1] = idx[j]; j--; }
      idx[j + 1] = k;
    }
    for (int i = 0; i < keep; i++) { out[o++] = s->items[first + idx[i]][0]; out[o++] = s->items[first + idx[i]][1]; }
  }
}

/* Gather task layer after physics: ant_gather_env.py:84-119 / gather_base.py:80-109.
 * `base` = robot obs (ant: 26 = state[0], state[3:]; point: 8).  replay != NULL replays uniforms. */
static void gather_task(hrlo_env* E, int e, env_state* s, const real* base, int nbase, real z_for_alive, int can_die,
                        real yaw, const double* replay, int* replay_used, const int* touched, real* obs, real* rew, int* done,
                        real* info) {
  const hrl_config* cfg = &E->cfg;
  int n = cfg->n_food + cfg->n_poison;
  real food_rew = 0;
  int used = 0;
  if (cfg->robot_coll_dist > 0) {
    for (int i = 0; i < n; i++) {
      real dx = s->items[i][0] - s->pos[0], dy = s->items[i][1] - s->pos[1];
      real d2 = dx * dx + dy * dy;
      if (d2 < (real)cfg->robot_coll_dist) { /* :90 (squared distance vs unsquared threshold) */
        food_rew += (i < cfg->n_food) ? 1 : -1; /* gather_scene.py:95-114 */
        if (cfg->respawn) {
          real o[2];
          int k = random_on_plane(cfg, s->pos[0], s->pos[1], cfg->seed, (uint32_t)(cfg->env_index_offset + e),
                                  STREAM_ITEM, (uint32_t)s->steps_total, i, replay ? replay + used : NULL, o);
          used += 2 * k;
          s->items[i][0] = o[0]; s->items[i][1] = o[1];
        } else { s->items[i][0] = 100; s->items[i][1] = 0; } /* fake_kill_pos gather_scene.py:13 */
      }
    }
  }
  if (replay_used) *replay_used = used;
  for (int i = 0; i < nbase; i++) obs[i] = base[i];
  food_obs(cfg, s, yaw, obs + nbase);
  int alive = can_die ? (z_for_alive > (real)0.26) : 1; /* Ant.alive_bonus; point_bot.py:73-74 */
  *done = !alive;
  if (!all_finite(obs, nbase + food_obs_dim(cfg))) *done = 1; /* :101-103 */
  if (!(cfg->robot_coll_dist > 0) && touched) {
    /* ant_gather_env.py:113-116: AFTER the observation is built, one reward_collision per contact POINT of the
     * robot with a cube (a cube touched by two spheres pays twice; it is respawned / parked once) */
    for (int i = 0; i < n; i++) {
      if (!touched[i]) continue;
      food_rew += (real)touched[i] * ((i < cfg->n_food) ? 1 : -1);
      if (cfg->respawn) {
        real o[2];
        int k = random_on_plane(cfg, s->pos[0], s->pos[1], cfg->seed, (uint32_t)(cfg->env_index_offset + e),
                                STREAM_ITEM, (uint32_t)s->steps_total, i, replay ? replay + used : NULL, o);
        used += 2 * k;
        s->items[i][0] = o[0]; s->items[i][1] = o[1];
      } else { s->items[i][0] = 100; s->items[i][1] = 0; }
    }
    if (replay_used) *replay_used = used;
  }
  real dead_rew = alive ? 0 : (real)cfg->dying_cost;
  *rew = food_rew + dead_rew;
  info[0] = food_rew; info[1] = dead_rew;
}

static void place_items(hrlo_env* E, int e, env_state* s) {
  const hrl_config* cfg = &E->cfg;
  /* gather_scene.py:38-50: every item re-randomised avoiding (0,0) */
  for (int i = 0; i < cfg->n_food + cfg->n_poison; i++) {
    real o[2];
    random_on_plane(cfg, 0, 0, cfg->seed, (uint32_t)(cfg->env_index_offset + e), STREAM_ITEM_RESET,
                    (uint32_t)s->episode, i, NULL, o);
    s->items[i][0] = o[0]; s->items[i][1] = o[1];
  }
}

/* Flagrun goal j of episode ep: ant_flagrun_env.py:71-78, stream shared by all envs (:39). */
static void flag_goal(const hrl_config* cfg, int episode, int j, int gen, real g[2]) {
  for (uint32_t attempt = 0;; attempt++) {
    real u[4]; /* gen = create_targets() calls so far: each call draws fresh goals from the stream (:91-96) */
    rng_u4(cfg->flag_seed, (uint32_t)gen, STREAM_FLAG, attempt, (uint32_t)(episode * 128 + j), u);
    real half = (real)cfg->flag_size / 2;
    g[0] = -half + 2 * half * u[0]; g[1] = -half + 2 * half * u[1];
    if (R_SQRT(g[0] * g[0] + g[1] * g[1]) < (real)0.5 && attempt + 1 < MAX_PLACE_ATTEMPTS) continue;
    return;
  }
}

static void joint_noise(hrlo_env* E, int e, env_state* s) {
  /* WalkerBase.robot_specific_reset: q ~ U(-0.1, 0.1), qd = 0.  Maze/Flagrun call it twice
     (ant_maze_bullet_env.py:111,118; ant_flagrun_env.py:141,143): only the last call survives,
     and RNG streams are not reproduced anyway, so one draw per episode. */
  const hrl_config* cfg = &E->cfg;
  real u[8];
  rng_u4(cfg->seed, (uint32_t)(cfg->env_index_offset + e), STREAM_JOINT, (uint32_t)s->episode, 0, u);
  rng_u4(cfg->seed, (uint32_t)(cfg->env_index_offset + e), STREAM_JOINT, (uint32_t)s->episode, 1, u + 4);
  for (int j = 0; j < 8; j++) { s->q[j] = (real)-0.1 + (real)0.2 * u[j]; s->qd[j] = 0; }
}

static int is_ant(int kind) { return kind != HRL_POINT_GATHER; }

int hrlo_scene_bounds(const hrl_config* cfg, double* b);
void hrlo_maze_target_sensor(int n_bins, double span, double range, int n_box, const double* box_bounds, double rx,
                             double ry, double yaw, double tx, double ty, double wtd, double* readings);

static void lidar(const hrl_config* cfg, const env_state* s, real yaw, real* out) {
  double bounds[7 * 4], w[HRL_MAX_BINS];
  int nl = hrlo_scene_bounds(cfg, bounds);
  hrlo_sense_walls_one(cfg->n_bins, snap_span(cfg->sensor_span), cfg->sensor_range, nl, bounds, (double)s->pos[0], (double)s->pos[1],
                       (double)yaw, w);
  for (int b = 0; b < cfg->n_bins; b++) out[b] = (real)w[b];
}

/* PointBot.calc_state (point_bot.py:48-67) for a general pose; walk target (0,0), initial_z 1 */
static void point_state_general(const real xyz[3], const real rpy[3], const real vel[3], real initial_z, real* o) {
  real a = R_ATAN2(0 - xyz[1], 0 - xyz[0]) - rpy[2];
  real cy = R_COS(-rpy[2]), sy = R_SIN(-rpy[2]);
  o[0] = xyz[2] - initial_z; o[1] = R_SIN(a); o[2] = R_COS(a);
  o[3] = (real)0.3 * (cy * vel[0] - sy * vel[1]); o[4] = (real)0.3 * (sy * vel[0] + cy * vel[1]); o[5] = (real)0.3 * vel[2];
  o[6] = rpy[0]; o[7] = rpy[1];
}
static void point_base_obs(const env_state* s, real* o) {
  /* the cube of this build only translates: roll = pitch = yaw = 0 */
  const real rpy[3] = {0, 0, 0};
  point_state_general(s->pos, rpy, s->vel, s->initial_z, o);
}

/* Observation assembled from a calc_state result `c` (ants) and the task state.  Used for
 * reset()'s return value, hrl_observe and the non-Gather step paths. */
static void compose_obs(const hrlo_env* E, const env_state* s, const calc_t* c, real* obs) {
  const hrl_config* cfg = &E->cfg;
  switch (cfg->env_kind) {
    case HRL_ANT_GATHER: { /* ant_gather_env.py:68-74 */
      obs[0] = c->obs28[0];
      for (int i = 3; i < 28; i++) obs[i - 2] = c->obs28[i];
      food_obs(cfg, s, c->rpy[2], obs + 26);
    } break;
    case HRL_ANT_MAZE: { /* ant_maze_bullet_env.py:63-75, :123-133 */
      obs[0] = c->obs28[0];
      for (int i = 3; i < 28; i++) obs[i - 2] = c->obs28[i];
      int nt = 2;
      if (cfg->sense_target) { /* get_target_sensor_obs :135-178 */
        double rd[HRL_MAX_BINS], bb[7 * 4];
        hrlo_scene_bounds(cfg, bb);
        hrlo_maze_target_sensor(cfg->n_bins, snap_span(cfg->sensor_span), (double)cfg->sensor_range, cfg->has_box ? 3 : 0, bb + 16,
                                (double)s->pos[0], (double)s->pos[1], (double)c->rpy[2], (double)s->target[0],
                                (double)s->target[1], (double)c->wtd, rd);
        nt = cfg->n_bins;
        for (int b = 0; b < nt; b++) obs[26 + b] = (real)rd[b];
      } else {
        real vx = s->target[0] - s->pos[0], vy = s->target[1] - s->pos[1];
        if (cfg->target_encoding == 0) { real nn = R_SQRT(vx * vx + vy * vy); obs[26] = vx / nn; obs[27] = vy / nn; }
        else { real a = R_ATAN2(vy, vx) - c->rpy[2]; obs[26] = R_SIN(a); obs[27] = R_COS(a); }
      }
      if (cfg->sense_walls) lidar(cfg, s, c->rpy[2], obs + 26 + nt);
    } break;
    case HRL_ANT_FLAGRUN:
      for (int i = 0; i < 28; i++) obs[i] = c->obs28[i];
      if (cfg->flag_use_sensor) lidar(cfg, s, c->rpy[2], obs + 28); /* ant_flagrun_env.py:122-130 (body_real_xyz = torso) */
      break;
    case HRL_ANT_MJ:
    case HRL_ANT_MAZE_MJ: { /* MjAnt.calc_state envs/MjAnt.py:17-25 */
      for (int i = 0; i < 3; i++) obs[i] = s->pos[i];
      for (int i = 0; i < 4; i++) obs[3 + i] = s->quat[i];
      for (int j = 0; j < 8; j++) obs[7 + j] = s->q[j];
      for (int i = 0; i < 3; i++) { obs[15 + i] = s->vel[i]; obs[18 + i] = s->ang[i]; }
      for (int j = 0; j < 8; j++) obs[21 + j] = s->qd[j];
      if (cfg->env_kind == HRL_ANT_MAZE_MJ) { /* ant_maze_mj_env.py:57-64 */
        lidar(cfg, s, c->rpy[2], obs + 29);
        for (int b = 0; b < cfg->n_bins; b++) { obs[29 + cfg->n_bins + b] = 0; obs[29 + 2 * cfg->n_bins + b] = 0; }
        obs[29 + 3 * cfg->n_bins] = (real)s->t * (real)0.001;
      }
    } break;
    case HRL_POINT_GATHER: { /* gather_base.py:67-72 */
      point_base_obs(s, obs);
      food_obs(cfg, s, 0, obs + 8);
    } break;
  }
}

static void write_obs(const hrlo_env* E, const env_state* s, real* obs) {
  calc_t c; memset(&c, 0, sizeof c);
  if (is_ant(E->cfg.env_kind)) ant_calc_state(E, s, &c);
  compose_obs(E, s, &c, obs);
}

/* Flagrun next_target (ant_flagrun_env.py:112-120): pops the next goal; the potential is
 * taken from the STALE cached walk_target_dist (quirk Q3), then calc_state refreshes it. */
static int flag_next_target(hrlo_env* E, int e, env_state* s, int ep, calc_t* c, int at_reset) {
  const hrl_config* cfg = &E->cfg;
  if (cfg->flag_max_targets < 1) {
    /* create_close_target (ant_flagrun_env.py:80-89): offset of magnitude U(tol, max_target_dist/2) per
     * axis with a random sign around the true torso xy, redrawn until strictly inside the world */
    real wb = (real)cfg->flag_size / 2, g0 = wb + 1, g1 = wb + 1;
    int inside = 0;
    for (uint32_t attempt = 0; attempt < MAX_CLOSE_ATTEMPTS && !inside; attempt++) {
      real u[4];
      rng_u4(cfg->flag_seed, (uint32_t)(cfg->env_index_offset + e), STREAM_FLAG_CLOSE, (uint32_t)s->steps_total,
             attempt * 2 + (uint32_t)(at_reset != 0), u);
      real lo = (real)cfg->tol, hi = (real)cfg->flag_max_target_dist / 2;
      g0 = (lo + (hi - lo) * u[0]) * (u[2] < (real)0.5 ? -1 : 1) + s->pos[0];
      g1 = (lo + (hi - lo) * u[1]) * (u[3] < (real)0.5 ? -1 : 1) + s->pos[1];
      inside = -wb < g0 && g0 < wb && -wb < g1 && g1 < wb;
    }
    if (!inside) { /* the reference redraws for ever; after 64 rejected draws the goal is pulled inside the world */
      real lim = wb - (real)1e-3f;
      g0 = g0 < -lim ? -lim : (g0 > lim ? lim : g0); g1 = g1 < -lim ? -lim : (g1 > lim ? lim : g1);
    }
    s->target[0] = g0; s->target[1] = g1;
  } else {
    if (s->goals_left <= 0) return 0; /* goals.pop() raises IndexError */
    s->goals_left--;
    if (E->replay_goals) { s->target[0] = (real)E->replay_goals[2 * s->goals_left]; s->target[1] = (real)E->replay_goals[2 * s->goals_left + 1]; }
    else flag_goal(cfg, ep, s->goals_left, s->goal_gen, s->target);
  }
  s->rewarded = 0;
  if (E->replay_stub_robot) { s->potential = -1; return 1; }
  s->potential = -s->wtd / (real)cfg->dt;
  ant_calc_state(E, s, c);
  s->wtd = c->wtd;
  return 1;
}

/* AntMazeBulletEnv.step after the inner walker step (ant_maze_bullet_env.py:84-96); s->t = steps taken
 * BEFORE this one, the reference's self.t (incremented at :78) is s->t + 1. */
static int maze_task(const hrl_config* cfg, const env_state* s, real inner, int done, real* rew) {
  *rew = inner * (real)cfg->inner_rew_weight;
  const int t = s->t + 1, last = (t == cfg->maze_max_steps - 1);
  if (s->wtd < (real)cfg->tol) {
    if (cfg->done_at_target || last) { *rew += 1; done = 1; }
  }
  if (last) done = 1;
  if (cfg->targ_dist_rew && done) *rew -= s->wtd; /* :93-94 */
  return done;
}

/* AntMazeMjEnv.step after the inner AntMjEnv step (ant_maze_mj_env.py:73-77). */
static int maze_mj_task(const hrl_config* cfg, const env_state* s, real inner, int done, real* rew) {
  *rew = inner * (real)cfg->inner_rew_weight;
  if (s->wtd < (real)cfg->tol) { *rew += 1; done = 1; }
  return done;
}

/* AntFlagrunBulletEnv.step after the inner walker step (ant_flagrun_env.py:167-202). */
static int flagrun_task(hrlo_env* E, int e, env_state* s, calc_t* c, real inner, int done, real* rew, int* switched) {
  const hrl_config* cfg = &E->cfg;
  real r = inner;
  s->since += 1;
  if (s->wtd < (real)cfg->tol) {
    if (!s->rewarded) { r += (real)cfg->goal_reach_rew; s->rewarded = 1; }
    if (cfg->flag_switch_on_collision) {
      if (flag_next_target(E, e, s, s->episode - 1, c, 0)) { s->since = 0; *switched = 1; }
      else done = 1; /* IndexError -> d = True (:193) */
    }
  }
  if (cfg->flag_timeout > 0 && cfg->flag_timeout <= s->since) {
    if (flag_next_target(E, e, s, s->episode - 1, c, 0)) { s->since = 0; *switched = 1; }
    else done = 1;
  }
  *rew = r;
  return done;
}

static void reset_env(hrlo_env* E, int e, real* obs) {
  hrl_config* cfg = &E->cfg;
  env_state* s = &E->s[e];
  int kind = cfg->env_kind;
  s->t = 0; s->ret = 0;
  for (int i = 0; i < 3; i++) { s->pos[i] = (real)cfg->start_pos[i]; s->vel[i] = 0; s->ang[i] = 0; }
  s->quat[0] = s->quat[1] = s->quat[2] = 0; s->quat[3] = 1;
  for (int k = 0; k < 4; k++) s->feet[k] = 0;
  if (kind == HRL_ANT_GATHER || kind == HRL_POINT_GATHER) place_items(E, e, s);
  if (is_ant(kind)) {
    joint_noise(E, e, s);
    s->initial_z = s->pos[2];
  } else {
    for (int j = 0; j < 8; j++) s->q[j] = s->qd[j] = 0;
    s->initial_z = 1; /* point_bot.py:18 */
  }
  if (kind == HRL_ANT_MAZE || kind == HRL_ANT_MAZE_MJ) {
    real u[4];
    rng_u4(cfg->seed, (uint32_t)(cfg->env_index_offset + e), STREAM_GOAL, (uint32_t)s->episode, 0, u);
    int idx = (int)(u[0] * (real)cfg->n_targets); /* rs.randint(0, len(targets)) ant_maze_bullet_env.py:110 */
    if (idx >= cfg->n_targets) idx = cfg->n_targets - 1;
    s->target[0] = (real)cfg->targets[idx][0]; s->target[1] = (real)cfg->targets[idx][1];
  }
  if (kind == HRL_ANT_GATHER || kind == HRL_POINT_GATHER) { s->target[0] = 0; s->target[1] = 0; }
  if (kind == HRL_ANT_MJ) { s->target[0] = 1000; s->target[1] = 0; }
  calc_t c; memset(&c, 0, sizeof c);
  if (kind == HRL_ANT_FLAGRUN) {
    /* ant_flagrun_env.py:141-153: calc_state with the STALE target, goals redrawn, next_target */
    if (s->episode == 0) { s->target[0] = 1000; s->target[1] = 0; } /* WalkerBase default walk target */
    ant_calc_state(E, s, &c);
    s->wtd = c.wtd;
    s->rewarded = 0;
    if (cfg->flag_manual_goals) { /* :150-153: goals.clear(), nothing drawn, the walk target stays; potential as in WalkerBase.reset */
      s->goals_left = 0; s->goal_gen = -1; /* the caller's first create_targets() draws generation 0, like the automatic mode */
      s->potential = -s->wtd / (real)cfg->dt;
    } else {
      s->goals_left = cfg->flag_max_targets; s->goal_gen = 0;
      flag_next_target(E, e, s, s->episode, &c, 1);
    }
  } else if (is_ant(kind)) {
    ant_calc_state(E, s, &c);
    s->wtd = c.wtd;
    s->potential = -s->wtd / (real)cfg->dt; /* Maze computes this before the target is set (B.5) */
  }
  s->episode++;
  if (obs) compose_obs(E, s, &c, obs);
}

/* sensor bound lines: sizeable_enclosed_scene.py:25-34 then maze_scene.py:15-21 */
int hrlo_scene_bounds(const hrl_config* cfg, double* b) {
  double x1 = cfg->world_size[0] / 2.0, y1 = cfg->world_size[1] / 2.0, x2 = -x1, y2 = -y1;
  double w[4][4] = {{x1, y1, x2, y1}, {x1, y1, x1, y2}, {x2, y2, x2, y1}, {x2, y2, x1, y2}};
  memcpy(b, w, sizeof w);
  int n = 4;
  if (cfg->has_box) {
    double bx1 = cfg->box_hi[0], by1 = cfg->box_hi[1], bx2 = cfg->box_lo[0], by2 = cfg->box_lo[1];
    double bb[3][4] = {{bx1, by1, bx1, by2}, {bx2, by2, bx2, by1}, {bx2, by2, bx1, by2}};
    memcpy(b + 16, bb, sizeof bb);
    n = 7;
  }
  return n;
}

/* one control step of env e.  Returns done. */
static int step_env(hrlo_env* E, int e, const float* act, real* obs, real* rew, real* info) {
  hrl_config* cfg = &E->cfg;
  env_state* s = &E->s[e];
  int kind = cfg->env_kind, done = 0, switched = 0;
  info[0] = info[1] = info[2] = info[3] = 0;
  *rew = 0;
  if (kind == HRL_POINT_GATHER) {
    /* point_bot.py:28-31: F = a/|a| * 500 (NaN when a == 0: kept, the finite guard ends the episode) */
    real ax = (real)act[0], ay = (real)act[1], nn = R_SQRT(ax * ax + ay * ay);
    real f[3] = {ax / nn * (real)cfg->torque_scale, ay / nn * (real)cfg->torque_scale, 0}, z3[3] = {0, 0, 0};
    for (int k = 0; k < cfg->substeps; k++) point_substep(E, s, (k == 0 || !cfg->torque_first_substep_only) ? f : z3);
    real base[8];
    point_base_obs(s, base);
    gather_task(E, e, s, base, 8, 1, 0, 0, NULL, NULL, NULL, obs, rew, &done, info);
  } else {
    real a[8], tau[8], zero[8] = {0};
    for (int j = 0; j < 8; j++) {
      real x = (real)act[j];
      a[j] = x > 1 ? 1 : (x < -1 ? -1 : x); /* WalkerBase.apply_action clip */
      tau[j] = (real)cfg->torque_scale * a[j];
    }
    int feet_ground[4] = {0, 0, 0, 0};
    int touched[HRL_MAX_ITEMS] = {0};
    for (int k = 0; k < cfg->substeps; k++) /* contact points as of the LAST internal step are what getContactPoints sees */
      if (ant_substep(E, s, (k == 0 || !cfg->torque_first_substep_only) ? tau : zero, feet_ground,
                      k == cfg->substeps - 1 ? touched : NULL)) { done = 1; break; }
    calc_t c;
    ant_calc_state(E, s, &c);
    s->wtd = c.wtd;
    real z = c.obs28[0] + s->initial_z; /* state[0] + initial_z (ant_gather_env.py:99) */
    if (kind == HRL_ANT_GATHER) {
      real base[26];
      base[0] = c.obs28[0];
      for (int i = 3; i < 28; i++) base[i - 2] = c.obs28[i];
      gather_task(E, e, s, base, 26, z, 1, c.rpy[2], NULL, NULL, touched, obs, rew, &done, info);
    } else {
      /* WalkerBaseBulletEnv.step [3P-MEM] SURVEY.md App. A.2 ; AntMjEnv.step envs/MjAnt.py:36-97 */
      int mj = (kind == HRL_ANT_MJ || kind == HRL_ANT_MAZE_MJ);
      int alive = mj ? (s->pos[2] > (real)0.26) : (z > (real)0.26);
      if (!alive) done = 1;
      real pot_old = s->potential;
      s->potential = -s->wtd / (real)cfg->dt;
      real progress = s->potential - pot_old;
      real elec = 0, sq = 0;
      for (int j = 0; j < 8; j++) { elec += R_FABS((real)act[j] * c.jspeed[j]); sq += (real)act[j] * (real)act[j]; }
      real electricity = mj ? 0 : (real)cfg->electricity_cost * (elec / 8) + (real)cfg->stall_torque_cost * (sq / 8);
      real limits = (real)cfg->joints_at_limit_cost * (real)c.joints_at_limit;
      real inner = (alive ? 1 : -1) + progress + electricity + limits;
      info[0] = inner;
      if (kind == HRL_ANT_MAZE || kind == HRL_ANT_MJ || kind == HRL_ANT_MAZE_MJ) {
        compose_obs(E, s, &c, obs); /* shows the PREVIOUS step's feet flags (quirk Q2) */
        if (!all_finite(mj ? obs : c.obs28, mj ? 29 : 28)) done = 1;
      } else if (!all_finite(c.obs28, 28)) done = 1;
      for (int k = 0; k < 4; k++) s->feet[k] = (real)feet_ground[k];
      if (kind == HRL_ANT_MAZE) {
        done = maze_task(cfg, s, inner, done, rew);
      } else if (kind == HRL_ANT_MAZE_MJ) {
        done = maze_mj_task(cfg, s, inner, done, rew);
      } else if (kind == HRL_ANT_FLAGRUN) {
        /* ant_flagrun_env.py:162-204 */
        done = flagrun_task(E, e, s, &c, inner, done, rew, &switched);
        compose_obs(E, s, &c, obs); /* the state after a goal switch (:120,190) + optional lidar */
        info[1] = (real)s->goals_left;
      } else {
        *rew = inner;
      }
    }
  }
  s->t++;
  s->steps_total++;
  /* gym TimeLimit */
  if (cfg->max_episode_steps > 0 && s->t >= cfg->max_episode_steps) { info[2] = done ? 0 : 1; done = 1; }
  info[2] += 2 * switched; /* bit 1: the walk target changed in this step (info['target'], ant_flagrun_env.py:188-191,199) */
  s->ret += *rew;
  if (done) { s->ret_sum += s->ret; s->ret = 0; }
  info[3] = (real)s->t;
  return done;
}

/* ======================================================================================
 * C API (mirrors include/hrl_b200.h with hrlo_ prefix and host `real` buffers)
 * ====================================================================================== */
int hrlo_real_size(void) { return (int)sizeof(real); }

int hrlo_default_config(int32_t kind, int32_t num_envs, hrl_config* c) {
  memset(c, 0, sizeof *c);
  c->env_kind = kind; c->num_envs = num_envs; c->seed = 0; c->max_episode_steps = 2000; c->auto_reset = 1;
  c->gravity = 9.8f; c->dt = 0.0165f; c->substeps = 4; c->solver_iters = 5;
  c->contact_erp = 0.9f; c->limit_erp = 0.2f; c->lin_damping = 0.04f; c->ang_damping = 0.04f;
  c->friction = 1.5f * 0.8f; c->limit_max_impulse = 100.f; c->max_coord_vel = 100.f; c->contact_margin = 0.02f;
  c->torque_scale = 250.f; c->torque_first_substep_only = 1;
  c->ground_z = 0.005f; c->has_walls = 1; c->has_box = 0;
  c->n_food = 8; c->n_poison = 8; c->n_bins = 10; c->sensor_range = 20.f; c->sensor_span = (float)PI_D;
  c->robot_coll_dist = 1.f; c->robot_object_spacing = 2.f; c->dying_cost = -10.f; c->respawn = 1; c->use_sensor = 1;
  c->tol = 1.5f; c->done_at_target = 1; c->inner_rew_weight = 0.f; c->target_encoding = 0; c->sense_walls = 1;
  c->flag_max_targets = 100; c->flag_timeout = 200; c->flag_size = 10.f; c->goal_reach_rew = 5000.f; c->flag_seed = 123;
  c->electricity_cost = -2.0f; c->stall_torque_cost = -0.1f; c->joints_at_limit_cost = -0.1f;
  c->sense_target = 0; c->maze_max_steps = -1; c->targ_dist_rew = 0;
  c->flag_use_sensor = 0; c->flag_switch_on_collision = 1; c->flag_max_target_dist = 0.f;
  c->item_contacts = 0; c->item_friction = 1.5f * 0.5f; c->item_half = 0.125f; c->item_z = 0.1f;
  switch (kind) {
    case HRL_ANT_GATHER:
      c->world_size[0] = c->world_size[1] = 15; c->start_pos[2] = 0.75f;
      c->item_contacts = 1; /* the food / poison cubes are real static colliders in the reference (gather_scene.py:66) */
      break;
    case HRL_POINT_GATHER:
      c->world_size[0] = c->world_size[1] = 15; c->start_pos[2] = 0.5f; c->n_bins = 5;
      c->friction = 0.1f * 0.8f; c->torque_scale = 500.f; break;
    case HRL_ANT_MAZE:
    case HRL_ANT_MAZE_MJ: {
      c->world_size[0] = 10; c->world_size[1] = 18; c->has_box = 1;
      c->box_lo[0] = -5; c->box_lo[1] = -2; c->box_lo[2] = 0; c->box_hi[0] = 1; c->box_hi[1] = 2; c->box_hi[2] = 2;
      c->start_pos[0] = -2; c->start_pos[1] = -5; c->start_pos[2] = 0.25f;
      c->sensor_range = 5.f; c->sensor_span = (float)(2 * PI_D);
      /* quirk Q1: floor (0,0) + last wall (-size_x/2, 0) + obstacle (-2,0) */
      c->n_scene_parts = 3; c->scene_parts_sum[0] = -7; c->scene_parts_sum[1] = 0;
      if (kind == HRL_ANT_MAZE) {
        const float t[4][2] = {{2, -3}, {2, 0}, {2, 3}, {-2, 4}};
        c->n_targets = 4; memcpy(c->targets, t, sizeof t);
      } else {
        const float t[5][2] = {{2, -4}, {2, 0}, {2, 4}, {0, 4}, {-2, 4}};
        c->n_targets = 5; memcpy(c->targets, t, sizeof t);
      }
    } break;
    case HRL_ANT_FLAGRUN:
      c->world_size[0] = c->world_size[1] = 12; c->start_pos[2] = 0.25f; c->tol = 0.5f;
      c->n_scene_parts = 2; c->scene_parts_sum[0] = -6; c->scene_parts_sum[1] = 0;
      c->electricity_cost = 0; c->stall_torque_cost = 0; c->joints_at_limit_cost = 0; /* ant_flagrun_env.py:133-135 */
      c->n_bins = 8; c->sensor_span = (float)PI_D; c->sensor_range = 4.f;            /* :15 (read when use_sensor) */
      break;
    case HRL_ANT_MJ:
      c->world_size[0] = c->world_size[1] = 50; c->has_walls = 0; c->ground_z = 0.f; c->start_pos[2] = 0.75f; break;
    default: return HRL_E_INVALID;
  }
  return HRL_OK;
}

int hrlo_obs_dim(const hrl_config* c) {
  switch (c->env_kind) {
    case HRL_ANT_GATHER: return 26 + food_obs_dim(c);
    case HRL_ANT_MAZE: return 26 + (c->sense_target ? c->n_bins : 2) + (c->sense_walls ? c->n_bins : 0);
    case HRL_ANT_FLAGRUN: return 28 + (c->flag_use_sensor ? c->n_bins : 0);
    case HRL_ANT_MJ: return 29;
    case HRL_ANT_MAZE_MJ: return 29 + 3 * c->n_bins + 1;
    case HRL_POINT_GATHER: return 8 + food_obs_dim(c);
  }
  return -1;
}
int hrlo_act_dim(const hrl_config* c) { return c->env_kind == HRL_POINT_GATHER ? 2 : 8; }

int hrlo_create(const hrl_config* cfg, hrlo_env** out) {
  if (!cfg || cfg->num_envs <= 0 || hrlo_obs_dim(cfg) < 0 || cfg->n_bins > HRL_MAX_BINS || cfg->n_food > 8 || cfg->n_poison > 8)
    return HRL_E_INVALID;
  hrlo_env* E = (hrlo_env*)calloc(1, sizeof *E);
  E->cfg = *cfg;
  ant_model_init(&E->model);
  E->s = (env_state*)calloc((size_t)cfg->num_envs, sizeof(env_state));
  for (int e = 0; e < cfg->num_envs; e++) E->s[e].quat[3] = 1;
  *out = E;
  return HRL_OK;
}
int hrlo_destroy(hrlo_env* E) { if (E) { free(E->s); free(E); } return HRL_OK; }

int hrlo_reset(hrlo_env* E, const uint8_t* mask, real* obs) {
  int D = hrlo_obs_dim(&E->cfg);
  for (int e = 0; e < E->cfg.num_envs; e++)
    if (!mask || mask[e]) reset_env(E, e, obs ? obs + (size_t)e * D : NULL);
  return HRL_OK;
}

/* envs [e0, e1) only: lets the caller spread shards over host threads */
int hrlo_step_range(hrlo_env* E, int e0, int e1, const float* actions, real* obs, real* rew, uint8_t* done, real* info,
                    real* terminal_obs) {
  int D = hrlo_obs_dim(&E->cfg), A = hrlo_act_dim(&E->cfg);
  for (int e = e0; e < e1; e++) {
    real inf[4];
    int d = step_env(E, e, actions + (size_t)e * A, obs + (size_t)e * D, &rew[e], inf);
    done[e] = (uint8_t)d;
    if (info) memcpy(info + 4 * (size_t)e, inf, sizeof inf);
    if (d && E->cfg.auto_reset) {
      if (terminal_obs) memcpy(terminal_obs + (size_t)e * D, obs + (size_t)e * D, sizeof(real) * (size_t)D);
      reset_env(E, e, obs + (size_t)e * D);
    }
  }
  return HRL_OK;
}
int hrlo_step(hrlo_env* E, const float* actions, real* obs, real* rew, uint8_t* done, real* info, real* terminal_obs) {
  return hrlo_step_range(E, 0, E->cfg.num_envs, actions, obs, rew, done, info, terminal_obs);
}

int hrlo_observe(hrlo_env* E, real* obs) {
  int D = hrlo_obs_dim(&E->cfg);
  for (int e = 0; e < E->cfg.num_envs; e++) write_obs(E, &E->s[e], obs + (size_t)e * D);
  return HRL_OK;
}

int hrlo_substeps(hrlo_env* E, const float* actions, int n_sub) {
  int A = hrlo_act_dim(&E->cfg);
  for (int e = 0; e < E->cfg.num_envs; e++) {
    env_state* s = &E->s[e];
    if (E->cfg.env_kind == HRL_POINT_GATHER) {
      real ax = actions[e * A], ay = actions[e * A + 1], nn = R_SQRT(ax * ax + ay * ay);
      real f[3] = {ax / nn * (real)E->cfg.torque_scale, ay / nn * (real)E->cfg.torque_scale, 0}, z3[3] = {0, 0, 0};
      for (int k = 0; k < n_sub; k++) point_substep(E, s, (k == 0 || !E->cfg.torque_first_substep_only) ? f : z3);
    } else {
      real tau[8], zero[8] = {0}; int fg[4];
      for (int j = 0; j < 8; j++) { real x = actions[e * A + j]; x = x > 1 ? 1 : (x < -1 ? -1 : x); tau[j] = (real)E->cfg.torque_scale * x; }
      for (int k = 0; k < n_sub; k++) ant_substep(E, s, (k == 0 || !E->cfg.torque_first_substep_only) ? tau : zero, fg, NULL);
    }
  }
  return HRL_OK;
}

int hrlo_get_state(hrlo_env* E, real* f, int32_t* iv) {
  for (int e = 0; e < E->cfg.num_envs; e++) {
    env_state* s = &E->s[e];
    real* o = f + (size_t)e * HRL_STATE_F;
    memset(o, 0, sizeof(real) * HRL_STATE_F);
    for (int i = 0; i < 3; i++) { o[HRL_SF_POS + i] = s->pos[i]; o[HRL_SF_LINVEL + i] = s->vel[i]; o[HRL_SF_ANGVEL + i] = s->ang[i]; }
    for (int i = 0; i < 4; i++) { o[HRL_SF_QUAT + i] = s->quat[i]; o[HRL_SF_FEET + i] = s->feet[i]; }
    for (int j = 0; j < 8; j++) { o[HRL_SF_Q + j] = s->q[j]; o[HRL_SF_QD + j] = s->qd[j]; }
    o[HRL_SF_INITIAL_Z] = s->initial_z; o[HRL_SF_POTENTIAL] = s->potential;
    o[HRL_SF_TARGET] = s->target[0]; o[HRL_SF_TARGET + 1] = s->target[1]; o[HRL_SF_WTD] = s->wtd;
    o[HRL_SF_RETURN] = s->ret; o[HRL_SF_RETURN_SUM] = s->ret_sum;
    for (int i = 0; i < HRL_MAX_ITEMS; i++) { o[HRL_SF_ITEMS + 2 * i] = s->items[i][0]; o[HRL_SF_ITEMS + 2 * i + 1] = s->items[i][1]; }
    int32_t* q = iv + (size_t)e * HRL_STATE_I;
    memset(q, 0, sizeof(int32_t) * HRL_STATE_I);
    q[HRL_SI_T] = s->t; q[HRL_SI_EPISODE] = s->episode; q[HRL_SI_STEPS] = s->steps_total;
    q[HRL_SI_GOALS_LEFT] = s->goals_left; q[HRL_SI_SINCE] = s->since; q[HRL_SI_REWARDED] = s->rewarded;
    q[HRL_SI_GOAL_GEN] = s->goal_gen;
  }
  return HRL_OK;
}
int hrlo_set_state(hrlo_env* E, const real* f, const int32_t* iv) {
  for (int e = 0; e < E->cfg.num_envs; e++) {
    env_state* s = &E->s[e];
    const real* o = f + (size_t)e * HRL_STATE_F;
    for (int i = 0; i < 3; i++) { s->pos[i] = o[HRL_SF_POS + i]; s->vel[i] = o[HRL_SF_LINVEL + i]; s->ang[i] = o[HRL_SF_ANGVEL + i]; }
    for (int i = 0; i < 4; i++) { s->quat[i] = o[HRL_SF_QUAT + i]; s->feet[i] = o[HRL_SF_FEET + i]; }
    for (int j = 0; j < 8; j++) { s->q[j] = o[HRL_SF_Q + j]; s->qd[j] = o[HRL_SF_QD + j]; }
    s->initial_z = o[HRL_SF_INITIAL_Z]; s->potential = o[HRL_SF_POTENTIAL];
    s->target[0] = o[HRL_SF_TARGET]; s->target[1] = o[HRL_SF_TARGET + 1]; s->wtd = o[HRL_SF_WTD];
    s->ret = o[HRL_SF_RETURN]; s->ret_sum = o[HRL_SF_RETURN_SUM];
    for (int i = 0; i < HRL_MAX_ITEMS; i++) { s->items[i][0] = o[HRL_SF_ITEMS + 2 * i]; s->items[i][1] = o[HRL_SF_ITEMS + 2 * i + 1]; }
    const int32_t* q = iv + (size_t)e * HRL_STATE_I;
    s->t = q[HRL_SI_T]; s->episode = q[HRL_SI_EPISODE]; s->steps_total = q[HRL_SI_STEPS];
    s->goals_left = q[HRL_SI_GOALS_LEFT]; s->since = q[HRL_SI_SINCE]; s->rewarded = q[HRL_SI_REWARDED];
    s->goal_gen = q[HRL_SI_GOAL_GEN];
  }
  return HRL_OK;
}

/* ---- test hooks for the golden vectors ------------------------------------------------ */
/* Gather task layer with injected post-physics robot state (tests/golden/gather_step.npz) */
int hrlo_gather_task_replay(const hrl_config* cfg, const double* base, int nbase, const double xyz[3], double yaw,
                            int can_die, double* items_xy, const double* uniforms, double* obs, double* rew_done_info,
                            int* used) {
  hrlo_env* E;
  hrl_config c = *cfg; c.num_envs = 1;
  if (hrlo_create(&c, &E)) return HRL_E_INVALID;
  env_state* s = &E->s[0];
  for (int i = 0; i < 3; i++) s->pos[i] = (real)xyz[i];
  for (int i = 0; i < HRL_MAX_ITEMS; i++) { s->items[i][0] = (real)items_xy[2 * i]; s->items[i][1] = (real)items_xy[2 * i + 1]; }
  real b[32], o[64], rew, info[4]; int done;
  for (int i = 0; i < nbase; i++) b[i] = (real)base[i];
  gather_task(E, 0, s, b, nbase, (real)xyz[2], can_die, (real)yaw, uniforms, used, NULL, o, &rew, &done, info);
  for (int i = 0; i < nbase + food_obs_dim(&c); i++) obs[i] = (double)o[i];
  for (int i = 0; i < HRL_MAX_ITEMS; i++) { items_xy[2 * i] = (double)s->items[i][0]; items_xy[2 * i + 1] = (double)s->items[i][1]; }
  rew_done_info[0] = (double)rew; rew_done_info[1] = done; rew_done_info[2] = (double)info[0]; rew_done_info[3] = (double)info[1];
  hrlo_destroy(E);
  return HRL_OK;
}

/* Maze task layer with an injected inner walker step (tests/golden/maze_step.npz):
 * obs28 = the walker's 28-d state, xy/yaw = true torso pose, wtd = robot.walk_target_dist */
int hrlo_maze_task_replay(const hrl_config* cfg, const double* obs28, const double xy[2], double yaw, double inner_rew,
                          int inner_done, double wtd, const double target[2], int t_before, double* obs, double* rew_done) {
  hrlo_env* E;
  hrl_config c = *cfg; c.num_envs = 1;
  if (hrlo_create(&c, &E)) return HRL_E_INVALID;
  env_state* s = &E->s[0];
  calc_t cs; memset(&cs, 0, sizeof cs);
  for (int i = 0; i < 28; i++) cs.obs28[i] = (real)obs28[i];
  cs.rpy[2] = (real)yaw; cs.wtd = (real)wtd;
  s->pos[0] = (real)xy[0]; s->pos[1] = (real)xy[1]; s->target[0] = (real)target[0]; s->target[1] = (real)target[1];
  s->wtd = (real)wtd; s->t = t_before;
  real o[64], rew;
  compose_obs(E, s, &cs, o);
  int done = maze_task(&c, s, (real)inner_rew, inner_done, &rew);
  for (int i = 0; i < hrlo_obs_dim(&c); i++) obs[i] = (double)o[i];
  rew_done[0] = (double)rew; rew_done[1] = done;
  hrlo_destroy(E);
  return HRL_OK;
}

/* AntMazeMjEnv.step / _get_obs with a stub inner AntMjEnv step (tests/golden/maze_mj_step.npz): obs29 = MjAnt's
 * observation (its first two entries are the xy the wall lidar is cast from, ant_maze_mj_env.py:58-59), yaw = the
 * torso yaw, wtd = robot.walk_target_dist, t_before = self.t on entry (the observation carries t_before * 0.001, :64). */
int hrlo_maze_mj_task_replay(const hrl_config* cfg, const double* obs29, double yaw, double inner_rew, int inner_done, double wtd,
                             int t_before, double* obs, double* rew_done) {
  hrlo_env* E;
  hrl_config c = *cfg; c.num_envs = 1;
  if (hrlo_create(&c, &E)) return HRL_E_INVALID;
  env_state* s = &E->s[0];
  calc_t cs; memset(&cs, 0, sizeof cs);
  for (int i = 0; i < 3; i++) { s->pos[i] = (real)obs29[i]; s->vel[i] = (real)obs29[15 + i]; s->ang[i] = (real)obs29[18 + i]; }
  for (int i = 0; i < 4; i++) s->quat[i] = (real)obs29[3 + i];
  for (int j = 0; j < 8; j++) { s->q[j] = (real)obs29[7 + j]; s->qd[j] = (real)obs29[21 + j]; }
  cs.rpy[2] = (real)yaw; cs.wtd = (real)wtd;
  s->wtd = (real)wtd; s->t = t_before;
  real o[64], rew;
  compose_obs(E, s, &cs, o);
  int done = maze_mj_task(&c, s, (real)inner_rew, inner_done, &rew);
  for (int i = 0; i < hrlo_obs_dim(&c); i++) obs[i] = (double)o[i];
  rew_done[0] = (double)rew; rew_done[1] = done;
  hrlo_destroy(E);
  return HRL_OK;
}

/* Flagrun step sequence with a scripted walk_target_dist / inner reward and an injected goal list
 * (tests/golden/flagrun_step.npz).  Returns the number of steps executed (stops after done). */
int hrlo_flagrun_replay(const hrl_config* cfg, const double* goals, int n_steps, const double* wtd, const double* inner_r,
                        double* rew, int* done_out, double* target, int* since, int* rewarded) {
  hrlo_env* E;
  hrl_config c = *cfg; c.num_envs = 1;
  if (hrlo_create(&c, &E)) return HRL_E_INVALID;
  env_state* s = &E->s[0];
  E->replay_goals = goals; E->replay_stub_robot = 1;
  s->episode = 1; s->goals_left = c.flag_max_targets; s->since = 0; s->rewarded = 0; s->wtd = 1;
  calc_t cs; memset(&cs, 0, sizeof cs);
  flag_next_target(E, 0, s, 0, &cs, 1);
  int i = 0;
  for (; i < n_steps; i++) {
    s->wtd = (real)wtd[i];
    real r;
    int sw = 0;
    int d = flagrun_task(E, 0, s, &cs, (real)inner_r[i], 0, &r, &sw);
    rew[i] = (double)r; done_out[i] = d; target[2 * i] = (double)s->target[0]; target[2 * i + 1] = (double)s->target[1];
    since[i] = s->since; rewarded[i] = s->rewarded;
    if (d) { i++; break; }
  }
  hrlo_destroy(E);
  return i;
}

/* PointBot.calc_state / apply_action and the AntMjEnv.step reward composition (tests/golden/robots.npz) */
void hrlo_point_state(const double xyz[3], const double rpy[3], const double vel[3], double* out8) {
  real a[3] = {(real)xyz[0], (real)xyz[1], (real)xyz[2]}, r[3] = {(real)rpy[0], (real)rpy[1], (real)rpy[2]};
  real v[3] = {(real)vel[0], (real)vel[1], (real)vel[2]}, o[8];
  point_state_general(a, r, v, 1, o);
  for (int i = 0; i < 8; i++) out8[i] = (double)o[i];
}
void hrlo_point_force(const hrl_config* cfg, const double act[2], double f[3]) {
  double nn = sqrt(act[0] * act[0] + act[1] * act[1]); /* point_bot.py:29 */
  f[0] = act[0] / nn * cfg->torque_scale; f[1] = act[1] / nn * cfg->torque_scale; f[2] = 0;
}
void hrlo_mj_reward(const hrl_config* cfg, double z, double pot_old, double pot_new, int joints_at_limit, double* rew_done) {
  /* envs/MjAnt.py:27-28,44,82-97: alive (z > 0.26) + progress + joints_at_limit_cost * count */
  int alive = z > 0.26;
  rew_done[0] = (alive ? 1.0 : -1.0) + (pot_new - pot_old) + (double)cfg->joints_at_limit_cost * joints_at_limit;
  rew_done[1] = !alive;
}

/* physics diagnostics for invariants tests: total mass, generalized mass matrix via impulse responses */
int hrlo_mass_matrix(hrlo_env* E, int e, real* Mout /*14x14 = inverse mass matrix*/) {
  kin_t K; real tau[8] = {0}, ud[NDOF];
  forward_kinematics(&E->model, &E->s[e], &K);
  if (aba(E, &E->s[e], &K, tau, ud)) return -1;
  for (int j = 0; j < NDOF; j++) {
    real f[NDOF] = {0}, dv[NDOF];
    f[j] = 1;
    impulse_response(E, &K, f, dv);
    for (int i = 0; i < NDOF; i++) Mout[i * NDOF + j] = dv[i];
  }
  return 0;
}
/* unconstrained generalized acceleration of env e (for energy / free-fall tests and CUDA cross-checks) */
int hrlo_free_accel(hrlo_env* E, int e, const real* tau, real* udot) {
  kin_t K;
  forward_kinematics(&E->model, &E->s[e], &K);
  return aba(E, &E->s[e], &K, tau, udot);
}
double hrlo_capsule_contacts(hrlo_env* E) { return E->n_capsule_contacts; }
/* test hook: closest interior point of the segment A..B to the box (capsule_interior_vs_box); out = Q(3), n(3), dist */
int hrlo_capsule_vs_box(const double A[3], const double B[3], double r, const float lo[3], const float hi[3], double out[7]) {
  v3 Q, n; real dist;
  if (!capsule_interior_vs_box(V3((real)A[0], (real)A[1], (real)A[2]), V3((real)B[0], (real)B[1], (real)B[2]), (real)r, lo, hi, &Q, &n, &dist)) return 0;
  for (int i = 0; i < 3; i++) { out[i] = (double)Q.v[i]; out[3 + i] = (double)n.v[i]; }
  out[6] = (double)dist;
  return 1;
}
void hrlo_stats(hrlo_env* E, double out[3]) { out[0] = E->n_contacts; out[1] = E->n_limit_rows; out[2] = E->n_substeps; }
void hrlo_rng_u4(uint64_t seed, uint32_t env, uint32_t stream, uint32_t draw, uint32_t sub, double* u) {
  real r[4]; rng_u4(seed, env, stream, draw, sub, r); for (int i = 0; i < 4; i++) u[i] = (double)r[i];
}
void hrlo_flag_goal(const hrl_config* cfg, int episode, int j, double* g) { real r[2]; flag_goal(cfg, episode, j, 0, r); g[0] = (double)r[0]; g[1] = (double)r[1]; }
