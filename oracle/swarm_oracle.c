/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference swarm transition, used
 *   (a) by tests/ as a fast second oracle (checked against the golden trajectories recorded from the
 *       unmodified reference, tests/golden/*.npz, and against oracle/swarm_oracle.py), and
 *   (b) by bench.py as the CPU baseline ("port"), OpenMP over environments.
 * The product path (gym-swarm_b200/) never links or calls this file.
 *
 * Parity status: PINNED by tests/test_oracle_golden.py::test_c_oracle_reproduces_reference.
 *
 * Every function cites the reference lines it follows:
 *   SDE = /root/reference/gym_swarm/envs/swarm_discrete_env.py
 * The loops are deliberately literal (O(N^2) collision check and reward, like the reference).
 * Randomness arrives as explicit tapes (see gym-swarm_b200/tapes.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ERR_SLAB_OVERFLOW 1u
#define ERR_PLACE_OVERFLOW 2u
#define ERR_PRED_OVERFLOW 4u
#define ERR_BAD_ACTION 8u
#define ERR_STEP_WHEN_DONE 16u

/* SDE:467-475 */
static const int MOVE_X[9] = {-1, -1, 0, 1, 1, 1, 0, -1, 0};
static const int MOVE_Y[9] = {0, -1, -1, -1, 0, 1, 1, 1, 0};

typedef struct {
  int E, N, S, att, rep;
  double eat;
  int K, PL, KP, R, auto_reset;
  int16_t *x, *y, *px, *py;
  int32_t* target;
  uint8_t *orient, *frozen;
  uint32_t *err, *reset_count;
  const int16_t *place, *ptape;
  double U[9][9]; /* cosine_distances(moves)/2 for every ordered action pair (off-diagonal value) */
} oracle_t;

/* SDE:45-57 / SDE:60-80 */
static int wrap1(int p, int S) {
  if (p > S - 1) return 0;
  if (p < 0) return S - 1;
  return p;
}

/* sklearn.metrics.pairwise.cosine_distances restated (SDE:346-347): normalize rows (zero rows stay
 * zero), 1 - <u,v>, clip to [0,2]; the /2 of SDE:347 is folded in. */
static void build_unalign_table(oracle_t* o) {
  for (int a = 0; a < 9; ++a)
    for (int b = 0; b < 9; ++b) {
      double ax = MOVE_X[a], ay = MOVE_Y[a], bx = MOVE_X[b], by = MOVE_Y[b];
      double na = sqrt(ax * ax + ay * ay), nb = sqrt(bx * bx + by * by);
      if (na == 0.0) na = 1.0;
      if (nb == 0.0) nb = 1.0;
      double d = 1.0 - ((ax / na) * (bx / nb) + (ay / na) * (by / nb));
      if (d < 0.0) d = 0.0;
      if (d > 2.0) d = 2.0;
      o->U[a][b] = d / 2.0;
    }
}

/* SDE:269-292: c[i] = number of agents on agent i's cell (the reference's ch_id repeats i c[i]-1 times).
 * Returns L = sum(c[i]-1) = len(ch_id). */
static int occupancy(const int* x, const int* y, int N, int* c) {
  int L = 0;
  for (int i = 0; i < N; ++i) {
    int n = 0;
    for (int j = 0; j < N; ++j) n += (x[i] == x[j]) & (y[i] == y[j]);
    c[i] = n;
    L += n - 1;
  }
  return L;
}

/* SDE:111-131 with stable argsort: lo/hi = lowest/highest index at the minimum Chebyshev distance. */
static int closest_target(int px, int py, const int* x, const int* y, int N, int S) {
  int dmin = 1 << 30, lo = 0, hi = 0;
  for (int i = 0; i < N; ++i) {
    int dx = abs(px - x[i]), dy = abs(py - y[i]);
    int d = dx > dy ? dx : dy;
    if (d < dmin) { dmin = d; lo = i; hi = i; }
    else if (d == dmin) hi = i;
  }
  /* id1 = lo, id2 = hi; if they differ: d1 = dmin, d2 = S - dmin, pick id1 iff d1 < d2 (SDE:124-129) */
  if (lo == hi) return lo;
  return (dmin < S - dmin) ? lo : hi;
}

/* SDE:197-221 + SDE:90-109 */
static void reset_env(oracle_t* o, int e, int* x, int* y, int* c) {
  const int N = o->N, S = o->S;
  uint32_t rc = o->reset_count[e]++;
  size_t slot = (size_t)(rc % (uint32_t)o->R) * o->E + e;
  const int16_t* tape = o->place + slot * o->PL;
  const int16_t* pt = o->ptape + slot * o->KP * 2;
  uint32_t err = 0;
  for (int i = 0; i < N; ++i) { x[i] = tape[i]; y[i] = tape[N + i]; }            /* SDE:202-203 */
  int cur = 2 * N;
  for (;;) {
    int L = occupancy(x, y, N, c);
    if (L == 0) break;
    if (cur + 2 * L > o->PL) { err |= ERR_PLACE_OVERFLOW; break; }
    int k = 0;                                                                    /* SDE:208-209 */
    for (int i = 0; i < N; ++i)
      for (int r = 0; r < c[i] - 1; ++r, ++k) { x[i] = tape[cur + k]; y[i] = tape[cur + L + k]; }
    cur += 2 * L;
  }
  int a = 0, px = 0, py = 0;
  for (;;) {                                                                      /* SDE:92-105 */
    if (a >= o->KP) { err |= ERR_PRED_OVERFLOW; break; }
    px = pt[2 * a]; py = pt[2 * a + 1]; ++a;
    int hit = 0;
    for (int i = 0; i < N; ++i) hit |= (x[i] == px) & (y[i] == py);
    if (!hit) break;
  }
  for (int i = 0; i < N; ++i) { o->x[(size_t)e * N + i] = (int16_t)x[i]; o->y[(size_t)e * N + i] = (int16_t)y[i]; }
  o->px[e] = (int16_t)px; o->py[e] = (int16_t)py;
  o->target[e] = closest_target(px, py, x, y, N, S);                               /* SDE:109 */
  o->orient[e] = 0;                                                                /* SDE:107 */
  o->err[e] |= err;
}

typedef struct {
  uint8_t* done; int32_t* eaten; double* glob; int64_t* cnt; int16_t *pred_prev, *pred_next; uint8_t* pred_orient;
  int32_t *step_target, *consumed; int16_t *term_x, *term_y; double* agent; int64_t* agent_cnt; uint8_t* eff;
} oracle_out;

/* SDE:223-267 for one env */
static void step_env(oracle_t* o, int e, const uint8_t* actions, int retarget, const uint8_t* slab, double w_att,
                     double w_rep, double w_ali, int vf, const oracle_out* out, int* scratch) {
  const int N = o->N, S = o->S;
  const size_t base = (size_t)e * N;
  int *xo = scratch, *yo = xo + N, *nx = yo + N, *ny = nx + N, *eff = ny + N, *c = eff + N;
  if (o->frozen[e]) {                                                             /* SDE:233-234 */
    o->err[e] |= ERR_STEP_WHEN_DONE;
    out->done[e] = 1; out->eaten[e] = -1;
    return;
  }
  uint32_t err = 0;
  for (int i = 0; i < N; ++i) {
    xo[i] = o->x[base + i]; yo[i] = o->y[base + i];
    int a = actions[base + i];
    if (a > 8) { err |= ERR_BAD_ACTION; a = 8; }
    eff[i] = a;
  }
  int cur = 0;
  for (;;) {                                                                      /* SDE:243-255 */
    for (int i = 0; i < N; ++i) {
      nx[i] = wrap1(xo[i] + MOVE_X[eff[i]], S);
      ny[i] = wrap1(yo[i] + MOVE_Y[eff[i]], S);
    }
    int L = occupancy(nx, ny, N, c);
    if (L == 0) break;
    if (cur + L > o->K) { err |= ERR_SLAB_OVERFLOW; break; }
    int k = 0;
    for (int i = 0; i < N; ++i)
      for (int r = 0; r < c[i] - 1; ++r, ++k) eff[i] = slab[(size_t)e * o->K + cur + k] & 7;   /* last draw wins */
    cur += L;
  }
  /* predator on the OLD agent positions: SDE:258-259, 133-165 */
  int px = o->px[e], py = o->py[e], target = o->target[e];
  out->pred_prev[2 * e] = (int16_t)px; out->pred_prev[2 * e + 1] = (int16_t)py;
  if (retarget) target = closest_target(px, py, xo, yo, N, S);
  int cdx = px - xo[target], cdy = py - yo[target];
  int mx = cdx >= 1 ? -1 : (cdx <= -1 ? 1 : 0), my = cdy >= 1 ? -1 : (cdy <= -1 ? 1 : 0);
  if ((cdx > cdy ? cdx : cdy) > S / 2) { mx = -mx; my = -my; }                      /* SDE:152-159 */
  int orient = 8;
  for (int a = 0; a < 9; ++a) if (MOVE_X[a] == mx && MOVE_Y[a] == my) orient = a;   /* SDE:161-163 */
  px = wrap1(px + mx, S); py = wrap1(py + my, S);
  out->pred_next[2 * e] = (int16_t)px; out->pred_next[2 * e + 1] = (int16_t)py;
  out->pred_orient[e] = (uint8_t)orient; out->step_target[e] = target; out->consumed[e] = cur;
  /* SDE:260 commit, SDE:307-327 termination */
  int eaten = -1;
  for (int i = 0; i < N; ++i) {
    o->x[base + i] = (int16_t)nx[i]; o->y[base + i] = (int16_t)ny[i];
    if (out->term_x) { out->term_x[base + i] = (int16_t)nx[i]; out->term_y[base + i] = (int16_t)ny[i]; }
    if (out->eff) out->eff[base + i] = (uint8_t)eff[i];
    if (eaten < 0 && nx[i] == px && ny[i] == py) eaten = i;
  }
  o->px[e] = (int16_t)px; o->py[e] = (int16_t)py; o->target[e] = target; o->orient[e] = (uint8_t)orient;
  double* g = out->glob + 5 * (size_t)e;
  g[0] = g[1] = g[2] = g[3] = g[4] = 0.0;
  out->cnt[2 * e] = out->cnt[2 * e + 1] = 0;
  if (out->agent) memset(out->agent + base * 5, 0, sizeof(double) * 5 * N);
  if (out->agent_cnt) memset(out->agent_cnt + base * 2, 0, sizeof(int64_t) * 2 * N);
  const int done = eaten >= 0;
  out->done[e] = (uint8_t)done; out->eaten[e] = eaten;
  if (done) {
    g[0] = o->eat; g[4] = o->eat;
    if (out->agent) { out->agent[(base + eaten) * 5 + 0] = o->eat; out->agent[(base + eaten) * 5 + 4] = o->eat; }
  } else {                                                                          /* SDE:330-379 */
    const double nn = (double)N * (double)(N - 1);
    const int att = o->att, rep = o->rep;
    long long tot_r1 = 0, tot_r2 = 0, tot_a = 0;
    double tot_u = 0.0;
    for (int i = 0; i < N; ++i) {
      long long r1 = 0, r2 = 0, aa = 0;
      double u = 0.0;
      for (int j = 0; j < N; ++j) {
        int dx = abs(nx[i] - nx[j]), dy = abs(ny[i] - ny[j]);
        int d1 = dx > dy ? dx : dy, d2 = S - d1;                                   /* SDE:332-333 */
        r1 += d1 < rep; r2 += d2 < rep;                                            /* SDE:336-337 */
        aa += ((d1 < att) == (d1 >= rep)) + ((d2 < att) == (d2 >= rep));           /* SDE:341-342 */
        if (j != i && !(vf >= 0 && d1 > vf)) u += o->U[eff[i]][eff[j]];            /* SDE:347-358 */
      }
      tot_r1 += r1; tot_r2 += r2; tot_a += aa; tot_u += u;
      if (out->agent) {                                                            /* SDE:371-379 */
        const long long rep_i = -(r1 + r2 - 1);
        double* ag = out->agent + (base + i) * 5;
        ag[2] = w_rep * (double)rep_i / nn;
        ag[1] = w_att * (double)aa / nn;
        ag[3] = w_ali * (-u) / nn;
        ag[4] = ag[2] + ag[1] + ag[3];
        if (out->agent_cnt) { out->agent_cnt[(base + i) * 2] = rep_i; out->agent_cnt[(base + i) * 2 + 1] = aa; }
      }
    }
    const long long rew_rep = -(tot_r1 - N + tot_r2);                               /* SDE:338 */
    g[2] = w_rep * (double)rew_rep / nn;                                           /* SDE:363-366 */
    g[1] = w_att * (double)tot_a / nn;
    g[3] = w_ali * (-tot_u) / nn;
    g[4] = g[2] + g[1] + g[3];
    out->cnt[2 * e] = rew_rep; out->cnt[2 * e + 1] = tot_a;
  }
  o->err[e] |= err;
  if (done) {
    if (o->auto_reset) reset_env(o, e, xo, yo, c);
    else o->frozen[e] = 1;
  }
}

/* ---------------------------------------------------------------- exported C API (ctypes) */

void* oracle_create(int E, int N, int S, int att, int rep, double eat, int K, int PL, int KP, int R, int auto_reset) {
  oracle_t* o = (oracle_t*)calloc(1, sizeof(oracle_t));
  o->E = E; o->N = N; o->S = S; o->att = att; o->rep = rep; o->eat = eat;
  o->K = K; o->PL = PL; o->KP = KP; o->R = R; o->auto_reset = auto_reset;
  o->x = (int16_t*)calloc((size_t)E * N, 2); o->y = (int16_t*)calloc((size_t)E * N, 2);
  o->px = (int16_t*)calloc(E, 2); o->py = (int16_t*)calloc(E, 2);
  o->target = (int32_t*)calloc(E, 4); o->orient = (uint8_t*)calloc(E, 1); o->frozen = (uint8_t*)calloc(E, 1);
  o->err = (uint32_t*)calloc(E, 4); o->reset_count = (uint32_t*)calloc(E, 4);
  build_unalign_table(o);
  return o;
}

void oracle_destroy(void* h) {
  oracle_t* o = (oracle_t*)h;
  free(o->x); free(o->y); free(o->px); free(o->py); free(o->target); free(o->orient); free(o->frozen);
  free(o->err); free(o->reset_count); free(o);
}

void oracle_set_reset_tapes(void* h, const int16_t* place, const int16_t* ptape) {
  oracle_t* o = (oracle_t*)h;
  o->place = place; o->ptape = ptape;
}

void oracle_reset(void* h) {
  oracle_t* o = (oracle_t*)h;
#pragma omp parallel
  {
    int* s = (int*)malloc(sizeof(int) * 3 * o->N);
#pragma omp for schedule(dynamic, 1)
    for (int e = 0; e < o->E; ++e) {
      o->reset_count[e] = 0; o->err[e] = 0; o->frozen[e] = 0;
      reset_env(o, e, s, s + o->N, s + 2 * o->N);
    }
    free(s);
  }
}

void oracle_step(void* h, const uint8_t* actions, const uint8_t* retarget, const uint8_t* slab, double w_att,
                 double w_rep, double w_ali, int vf, uint8_t* done, int32_t* eaten, double* glob, int64_t* cnt,
                 int16_t* pred_prev, int16_t* pred_next, uint8_t* pred_orient, int32_t* step_target, int32_t* consumed,
                 int16_t* term_x, int16_t* term_y, double* agent, int64_t* agent_cnt, uint8_t* eff) {
  oracle_t* o = (oracle_t*)h;
  oracle_out out = {done, eaten, glob, cnt, pred_prev, pred_next, pred_orient, step_target, consumed,
                    term_x, term_y, agent, agent_cnt, eff};
#pragma omp parallel
  {
    int* s = (int*)malloc(sizeof(int) * 6 * o->N);
#pragma omp for schedule(dynamic, 1)
    for (int e = 0; e < o->E; ++e)
      step_env(o, e, actions, retarget[e], slab, w_att, w_rep, w_ali, vf, &out, s);
    free(s);
  }
}

/* bench.py's CPU arms: use n host threads regardless of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1). */
int oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

void oracle_get_state(void* h, int16_t* x, int16_t* y, int16_t* px, int16_t* py, int32_t* target, uint32_t* err,
                      uint8_t* frozen) {
  oracle_t* o = (oracle_t*)h;
  memcpy(x, o->x, (size_t)o->E * o->N * 2); memcpy(y, o->y, (size_t)o->E * o->N * 2);
  memcpy(px, o->px, (size_t)o->E * 2); memcpy(py, o->py, (size_t)o->E * 2);
  memcpy(target, o->target, (size_t)o->E * 4); memcpy(err, o->err, (size_t)o->E * 4);
  memcpy(frozen, o->frozen, (size_t)o->E);
}
