/*
 * mgym_oracle.c -- scalar CPU restatement of the reference's classic-control
 * step/reset arithmetic.  TEST INFRASTRUCTURE ONLY (see mgym_oracle.h).
 *
 * Every env function cites the reference lines it follows
 * (paths relative to /root/reference).  All env arithmetic is IEEE binary32,
 * one rounding per operation, evaluated in the reference's operator order;
 * this file MUST be compiled with -ffp-contract=off.
 */
#include "mgym_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* Clone the exported hot entry points for FMA hardware so __builtin_fma inlines
 * to vfmadd (same result as libm's fma(), just faster).  -ffp-contract=off keeps
 * the f32 env arithmetic un-fused in both clones. */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define ORACLE_HOT __attribute__((target_clones("default", "fma")))
#else
#define ORACLE_HOT
#endif

/* ------------------------------------------------------------------------- */
/* glibc 2.39 sinf / cosf  (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c,       */
/* sincosf.h, s_sincosf_data.c; Szabolcs Nagy's implementation).  The FMA    */
/* placement is the one gcc chose for glibc's x86_64 multiarch __sinf_fma /  */
/* __cosf_fma objects in this image's libm.so.6, which is what Rust's        */
/* f32::sin / f32::cos (cartpole.rs:264-265, mountain_car.rs:302) call.      */
/* ------------------------------------------------------------------------- */

typedef struct {
  double sign[4];
  double hpi_inv, hpi;
  double c0, c1, c2, c3, c4;
  double s1, s2, s3;
} sincos_tab;

static const sincos_tab k_sincos[2] = {
    {{1.0, -1.0, -1.0, 1.0},
     0x1.45F306DC9C883p+23,
     0x1.921FB54442D18p0,
     0x1p0,
     -0x1.ffffffd0c621cp-2,
     0x1.55553e1068f19p-5,
     -0x1.6c087e89a359dp-10,
     0x1.99343027bf8c3p-16,
     -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7,
     -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0},
     0x1.45F306DC9C883p+23,
     0x1.921FB54442D18p0,
     -0x1p0,
     0x1.ffffffd0c621cp-2,
     -0x1.55553e1068f19p-5,
     0x1.6c087e89a359dp-10,
     -0x1.99343027bf8c3p-16,
     -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7,
     -0x1.994eb3774cf24p-13}};

/* 4/pi as a 192-bit integer, in overlapping 32-bit windows (__inv_pio4). */
static const uint32_t k_inv_pio4[24] = {
    0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};

static inline uint32_t f32_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline uint32_t abstop12(float x) { return (f32_bits(x) >> 20) & 0x7ff; }

/* sinf_poly, n even: sine polynomial; n odd: cosine polynomial. */
static inline float sincos_poly(double x, double x2, const sincos_tab *p, int n) {
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double s1 = __builtin_fma(x2, p->s3, p->s2);
    double x7 = x3 * x2;
    double s = __builtin_fma(x3, p->s1, x);
    return (float)__builtin_fma(s1, x7, s);
  } else {
    double x4 = x2 * x2;
    double c2 = __builtin_fma(x2, p->c4, p->c3);
    double c1 = __builtin_fma(x2, p->c1, p->c0);
    double x6 = x4 * x2;
    double c = __builtin_fma(x4, p->c2, c1);
    return (float)__builtin_fma(c2, x6, c);
  }
}

/* reduce_fast: |x| < 120.  n = round(x / (pi/2)), returns x - n*pi/2. */
static inline double reduce_fast(double x, int *np) {
  double r = x * k_sincos[0].hpi_inv;
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return __builtin_fma(-(double)n, k_sincos[0].hpi, x);
}

/* reduce_large: 120 <= |x| < inf, on the raw bits of x. */
static inline double reduce_large(uint32_t xi, int *np) {
  const uint32_t *arr = &k_inv_pio4[(xi >> 26) & 15];
  int shift = (xi >> 23) & 7;
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  res0 = (uint32_t)(xi * arr[0]);
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  double x = (double)(int64_t)res0;
  *np = (int)n;
  return x * 0x1.921FB54442D18p-62;
}

static inline float sincos_eval(float y, int want_cos) {
  double x = (double)y;
  int n;
  const sincos_tab *p = &k_sincos[0];
  uint32_t top = abstop12(y);
  if (top < 0x3f4) { /* |y| < pi/4 */
    if (top < 0x398) /* |y| < 2^-12 */
      return want_cos ? 1.0f : y;
    return sincos_poly(x, x * x, p, want_cos);
  } else if (top < 0x42f) { /* |y| < 120 */
    x = reduce_fast(x, &n);
    double s = p->sign[n & 3];
    if (n & 2) p = &k_sincos[1];
    return sincos_poly(x * s, x * x, p, n ^ want_cos);
  } else if (top < 0x7f8) {
    uint32_t xi = f32_bits(y);
    int sign = (int)(xi >> 31);
    x = reduce_large(xi, &n);
    double s = p->sign[(n + sign) & 3];
    if ((n + sign) & 2) p = &k_sincos[1];
    return sincos_poly(x * s, x * x, p, n ^ want_cos);
  }
  return y - y; /* inf/nan -> nan (__math_invalidf) */
}

static inline float ref_sinf(float x) { return sincos_eval(x, 0); }
static inline float ref_cosf(float x) { return sincos_eval(x, 1); }

ORACLE_HOT float oracle_sinf(float x) { return ref_sinf(x); }
ORACLE_HOT float oracle_cosf(float x) { return ref_cosf(x); }

/* Order-independent 64-bit digest of (x, sin x, cos x) over the bit patterns first, first+stride, ... (count of
 * them, both signs): lets the device be compared with this restatement over the WHOLE domain without moving
 * 2^32 results around (tests/test_gpu_trig_exhaustive.py; the device computes the same sum with atomics). */
static inline uint64_t trig_mix(uint32_t xb, uint32_t sb, uint32_t cb) {
  uint64_t h = ((uint64_t)sb << 32) | cb;
  h ^= (uint64_t)xb * 0x9E3779B97F4A7C15ull;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  return h;
}

typedef struct {
  uint32_t first, stride;
  uint64_t count, sum;
} trig_job;

ORACLE_HOT static void trig_checksum_run(trig_job *j) {
  uint64_t sum = 0;
  for (uint64_t i = 0; i < j->count; ++i) {
    const uint32_t mag = j->first + (uint32_t)(i * j->stride);
    for (uint32_t sign = 0; sign < 2; ++sign) {
      const uint32_t xb = mag | (sign << 31);
      float x;
      memcpy(&x, &xb, 4);
      const float sv = ref_sinf(x), cv = ref_cosf(x);
      sum += trig_mix(xb, f32_bits(sv), f32_bits(cv));
    }
  }
  j->sum = sum;
}
static void *trig_checksum_thread(void *arg) {
  trig_checksum_run((trig_job *)arg);
  return NULL;
}

uint64_t oracle_trig_checksum(uint32_t first_bits, uint64_t count, uint32_t stride, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  trig_job jobs[256];
  pthread_t th[256];
  const uint64_t per = (count + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    uint64_t a = per * t, b = a + per;
    if (a > count) a = count;
    if (b > count) b = count;
    jobs[t].first = first_bits + (uint32_t)(a * stride);
    jobs[t].stride = stride;
    jobs[t].count = b - a;
    jobs[t].sum = 0;
    pthread_create(&th[t], NULL, trig_checksum_thread, &jobs[t]);
  }
  uint64_t sum = 0;
  for (int t = 0; t < n_threads; ++t) {
    pthread_join(th[t], NULL);
    sum += jobs[t].sum;
  }
  return sum;
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123 constants)    */
/* ------------------------------------------------------------------------- */

static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  uint32_t n1 = (uint32_t)p1;
  uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  uint32_t n3 = (uint32_t)p0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, 16);
}

/* counter = (g.lo, g.hi, t.lo, t.hi[23:0] | tag<<24), key = seed */
static inline void philox_env(uint64_t seed, uint64_t g, uint64_t t, uint32_t tag, uint32_t out[4]) {
  uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)t,
                     ((uint32_t)(t >> 32) & 0x00FFFFFFu) | (tag << 24)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  oracle_philox4x32_10(ctr, key, out);
}

/* U[lo,hi) sampled in f64 then cast to f32 (cartpole.rs:240-241, mountain_car.rs:281-282) */
static inline float uniform_f64_to_f32(uint32_t w, double lo, double hi) {
  double u = (double)w * 0x1p-32;
  double v = u * (hi - lo);
  v = v + lo;
  return (float)v;
}

#define ORACLE_PI_D 3.14159265358979323846

/* ------------------------------------------------------------------------- */
/* env parameters, evaluated in f32 in the constructors' operator order       */
/* ------------------------------------------------------------------------- */

typedef struct {
  float gravity, masspole, total_mass, length, polemass_length, force_mag, tau;
  float x_threshold, theta_threshold;
} cartpole_params;

/* cartpole.rs:45-56 */
static inline cartpole_params cartpole_new(void) {
  cartpole_params p;
  float masscart = 1.0f;
  p.gravity = 9.8f;
  p.masspole = 0.1f;
  p.total_mass = p.masspole + masscart;
  p.length = 0.5f;
  p.polemass_length = p.masspole * p.length;
  p.force_mag = 10.0f;
  p.tau = 0.02f;
  /* 12.0 * 2.0 * PI / 360.0, left to right in f32 */
  float t = 12.0f * 2.0f;
  t = t * 3.14159274101257324f; /* std::f32::consts::PI */
  p.theta_threshold = t / 360.0f;
  p.x_threshold = 2.4f;
  return p;
}

typedef struct {
  float min_position, max_position, max_speed, goal_position, force, gravity;
} mountain_car_params;

/* mountain_car.rs:35-40 */
static inline mountain_car_params mountain_car_new(void) {
  mountain_car_params p = {-1.2f, 0.6f, 0.07f, 0.5f, 0.001f, 0.0025f};
  return p;
}

static inline float clampf(float x, float lo, float hi) { /* f32::clamp */
  if (x < lo) x = lo;
  if (x > hi) x = hi;
  return x;
}

static inline uint32_t sat_inc(uint32_t v) { return v == 0xFFFFFFFFu ? v : v + 1u; }

/* ------------------------------------------------------------------------- */
/* CartPoleV1::step -- cartpole.rs:251-348                                    */
/* ------------------------------------------------------------------------- */
static inline uint32_t cartpole_step(const oracle_config *cfg, float *st, uint32_t *steps,
                                     uint32_t *sbt, uint32_t action, float *reward) {
  const cartpole_params p = cartpole_new();
  float x = st[0], x_dot = st[1], theta = st[2], theta_dot = st[3]; /* :253-255 */
  float force = (action == 0) ? -p.force_mag : p.force_mag;           /* :258-262 */
  float costheta = ref_cosf(theta);                                   /* :264 */
  float sintheta = ref_sinf(theta);                                   /* :265 */

  /* :267-268  (force + pml*theta_dot*theta_dot*sintheta) / total_mass */
  float t0 = p.polemass_length * theta_dot;
  t0 = t0 * theta_dot;
  t0 = t0 * sintheta;
  float temp = (force + t0) / p.total_mass;
  /* :269-270 */
  float num = p.gravity * sintheta - costheta * temp;
  float d0 = p.masspole * costheta;
  d0 = d0 * costheta;
  d0 = d0 / p.total_mass;
  float four_thirds = 4.0f / 3.0f;
  float den = p.length * (four_thirds - d0);
  float thetaacc = num / den;
  /* :271 */
  float t1 = p.polemass_length * thetaacc;
  t1 = t1 * costheta;
  t1 = t1 / p.total_mass;
  float xacc = temp - t1;

  if (cfg->is_euler) { /* :273-277 */
    x = x + p.tau * x_dot;
    x_dot = x_dot + p.tau * xacc;
    theta = theta + p.tau * theta_dot;
    theta_dot = theta_dot + p.tau * thetaacc;
  } else { /* :278-283, verbatim (x is never updated; theta_dot is updated twice) */
    float half_tau = 0.5f * p.tau;
    x_dot = x_dot + half_tau * (xacc + temp);
    theta_dot = theta_dot + half_tau * (thetaacc + temp);
    theta = theta + (p.tau * theta_dot + (half_tau * p.tau) * thetaacc);
    theta_dot = theta_dot + half_tau * (thetaacc + temp);
  }
  st[0] = x; /* :285-290 */
  st[1] = x_dot;
  st[2] = theta;
  st[3] = theta_dot;

  int terminated = x < -p.x_threshold || x > p.x_threshold || theta < -p.theta_threshold ||
                   theta > p.theta_threshold; /* :291-294 */

  *steps = sat_inc(*steps); /* :296 */
  if (*steps >= 500) {      /* :297-306 early return */
    *sbt = 1;               /* Some(0) */
    *reward = 1.0f;
    return ORACLE_FLAG_TRUNCATED;
  }
  if (!terminated) { /* :310-318 */
    *reward = cfg->sutton_barto_reward ? 0.0f : 1.0f;
    return 0;
  } else if (*sbt == ORACLE_SBT_NONE) { /* :319-329 pole just fell */
    *sbt = 1;
    *reward = cfg->sutton_barto_reward ? -1.0f : 1.0f;
    return ORACLE_FLAG_TERMINATED;
  } else { /* :330-347 */
    *reward = cfg->sutton_barto_reward ? -1.0f : 0.0f;
    *sbt = sat_inc(*sbt);
    return ORACLE_FLAG_TERMINATED;
  }
}

/* ------------------------------------------------------------------------- */
/* MountainCarV0::step -- mountain_car.rs:293-330                             */
/* max_episode_steps > 0 adds Gymnasium's TimeLimit (not in the reference,    */
/* which hard-codes truncated=false, :328); 0 is the reference behaviour.     */
/* ------------------------------------------------------------------------- */
static inline uint32_t time_limit(const oracle_config *cfg, uint32_t *steps) {
  *steps = sat_inc(*steps);
  return (cfg->max_episode_steps > 0 && *steps >= (uint32_t)cfg->max_episode_steps)
             ? ORACLE_FLAG_TRUNCATED
             : 0u;
}

static inline uint32_t mountain_car_step(const oracle_config *cfg, float *st, uint32_t *steps,
                                         uint32_t action, float *reward) {
  const mountain_car_params p = mountain_car_new();
  float position = st[0], velocity = st[1]; /* :296-297 */
  /* :301-302  velocity += (a as f32 - 1.0)*force + cos(3.0*position)*(-gravity) */
  float a = ((float)action - 1.0f) * p.force;
  float b = ref_cosf(3.0f * position) * (-p.gravity);
  velocity = velocity + (a + b);
  velocity = clampf(velocity, -p.max_speed, p.max_speed);          /* :304 */
  position = position + velocity;                                  /* :306 */
  position = clampf(position, p.min_position, p.max_position);     /* :308 */
  if (position == p.min_position && velocity < 0.0f) velocity = 0.0f; /* :311-313 */
  st[0] = position; /* :315 */
  st[1] = velocity;
  int terminated = position >= p.goal_position && velocity >= cfg->goal_velocity; /* :318 */
  *reward = -1.0f;                                                                /* :319 */
  return (terminated ? ORACLE_FLAG_TERMINATED : 0u) | time_limit(cfg, steps);
}

/* ------------------------------------------------------------------------- */
/* MountainCarContinuous-v0 -- NOT IN THE REFERENCE (parity unpinned).        */
/* f32 restatement of Gymnasium continuous_mountain_car.py step().            */
/* ------------------------------------------------------------------------- */
static inline uint32_t mountain_car_continuous_step(const oracle_config *cfg, float *st,
                                                    uint32_t *steps, float action, float *reward) {
  const float min_position = -1.2f, max_position = 0.6f, max_speed = 0.07f;
  const float goal_position = 0.45f, power = 0.0015f;
  float position = st[0], velocity = st[1];
  float force = action;
  if (force < -1.0f) force = -1.0f; /* min(max(a, -1), 1) */
  if (force > 1.0f) force = 1.0f;
  velocity = velocity + (force * power - 0.0025f * ref_cosf(3.0f * position));
  if (velocity > max_speed) velocity = max_speed;
  if (velocity < -max_speed) velocity = -max_speed;
  position = position + velocity;
  if (position > max_position) position = max_position;
  if (position < min_position) position = min_position;
  if (position == min_position && velocity < 0.0f) velocity = 0.0f;
  int terminated = position >= goal_position && velocity >= cfg->goal_velocity;
  float r = terminated ? 100.0f : 0.0f;
  r = r - (action * action) * 0.1f;
  st[0] = position;
  st[1] = velocity;
  *reward = r;
  return (terminated ? ORACLE_FLAG_TERMINATED : 0u) | time_limit(cfg, steps);
}

/* ------------------------------------------------------------------------- */
/* Pendulum-v1 -- NOT IN THE REFERENCE (parity unpinned).                     */
/* f32 restatement of Gymnasium pendulum.py step(): g=10, m=1, l=1, dt=0.05.  */
/* ------------------------------------------------------------------------- */
#define ORACLE_PI_F 3.14159274101257324f
#define ORACLE_TWO_PI_F 6.28318548202514648f
#define ORACLE_HALF_PI_F 1.57079637050628662f

static inline float angle_normalize(float x) { /* ((x + pi) % (2 pi)) - pi, Python floored mod */
  float t = x + ORACLE_PI_F;
  float m = fmodf(t, ORACLE_TWO_PI_F);
  if (m != 0.0f) {
    if (m < 0.0f) m = m + ORACLE_TWO_PI_F;
  } else {
    m = 0.0f;
  }
  return m - ORACLE_PI_F;
}

static inline uint32_t pendulum_step(const oracle_config *cfg, float *st, uint32_t *steps,
                                     float action, float *reward) {
  const float dt = 0.05f, max_speed = 8.0f, max_torque = 2.0f;
  float th = st[0], thdot = st[1];
  float u = clampf(action, -max_torque, max_torque);
  float an = angle_normalize(th);
  float costs = an * an + 0.1f * (thdot * thdot);
  costs = costs + 0.001f * (u * u);
  /* 3g/(2l) = 15, 3/(m l^2) = 3 */
  float acc = 15.0f * ref_sinf(th) + 3.0f * u;
  float newthdot = thdot + acc * dt;
  newthdot = clampf(newthdot, -max_speed, max_speed);
  float newth = th + newthdot * dt;
  st[0] = newth;
  st[1] = newthdot;
  *reward = -costs;
  return time_limit(cfg, steps);
}

/* ------------------------------------------------------------------------- */
/* Acrobot-v1 -- NOT IN THE REFERENCE (parity unpinned).                      */
/* f32 restatement of Gymnasium acrobot.py ("book" dynamics, one RK4 step).   */
/* ------------------------------------------------------------------------- */
typedef struct {
  float m1lc1g; /* (m1*lc1 + m2*l1) * g */
  float m2lc2g; /* m2*lc2*g */
  float dt, dt2, dt6, max_vel_1, max_vel_2;
} acrobot_params;

static inline acrobot_params acrobot_new(void) {
  acrobot_params p;
  const float m1 = 1.0f, m2 = 1.0f, l1 = 1.0f, lc1 = 0.5f, lc2 = 0.5f, g = 9.8f;
  p.m1lc1g = (m1 * lc1 + m2 * l1) * g;
  p.m2lc2g = (m2 * lc2) * g;
  p.dt = 0.2f;
  p.dt2 = p.dt / 2.0f;
  p.dt6 = p.dt / 6.0f;
  p.max_vel_1 = 4.0f * ORACLE_PI_F;
  p.max_vel_2 = 9.0f * ORACLE_PI_F;
  return p;
}

static inline void acrobot_dsdt(const acrobot_params *p, const float s[4], float a, float d[4]) {
  float theta1 = s[0], theta2 = s[1], dtheta1 = s[2], dtheta2 = s[3];
  float c2 = ref_cosf(theta2), s2 = ref_sinf(theta2);
  /* d1 = m1*lc1^2 + m2*(l1^2 + lc2^2 + 2*l1*lc2*cos(theta2)) + I1 + I2 */
  float d1 = (0.25f + (1.25f + c2)) + 1.0f;
  d1 = d1 + 1.0f;
  /* d2 = m2*(lc2^2 + l1*lc2*cos(theta2)) + I2 */
  float d2 = (0.25f + 0.5f * c2) + 1.0f;
  float phi2 = p->m2lc2g * ref_cosf((theta1 + theta2) - ORACLE_HALF_PI_F);
  float phi1 = (-0.5f * (dtheta2 * dtheta2)) * s2 - ((dtheta2 * dtheta1) * s2);
  phi1 = phi1 + p->m1lc1g * ref_cosf(theta1 - ORACLE_HALF_PI_F);
  phi1 = phi1 + phi2;
  float num = a + (d2 / d1) * phi1;
  num = num - (0.5f * (dtheta1 * dtheta1)) * s2;
  num = num - phi2;
  float ddtheta2 = num / (1.25f - (d2 * d2) / d1);
  float ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
  d[0] = dtheta1;
  d[1] = dtheta2;
  d[2] = ddtheta1;
  d[3] = ddtheta2;
}

/* Gymnasium's wrap() loops `while x > M: x -= diff` without bound, which never ends for a huge or infinite
 * x (x - 2*pi == x).  From any state the step itself can produce, |x| < 4*pi and at most two passes run, so
 * the loops are capped at 4 passes: same values on every reachable state, and termination on garbage. */
#define ORACLE_WRAP_MAX_PASSES 4
static inline float wrapf(float x, float m, float M) {
  float diff = M - m;
  for (int i = 0; i < ORACLE_WRAP_MAX_PASSES && x > M; ++i) x = x - diff;
  for (int i = 0; i < ORACLE_WRAP_MAX_PASSES && x < m; ++i) x = x + diff;
  return x;
}

static inline float boundf(float x, float m, float M) { /* min(max(x, m), M) */
  float t = (m > x) ? m : x;
  return (M < t) ? M : t;
}

static inline uint32_t acrobot_step(const oracle_config *cfg, float *st, uint32_t *steps,
                                    uint32_t action, float *reward) {
  const acrobot_params p = acrobot_new();
  float torque = (float)action - 1.0f; /* AVAIL_TORQUE = [-1, 0, +1] */
  float k1[4], k2[4], k3[4], k4[4], y[4];
  acrobot_dsdt(&p, st, torque, k1);
  for (int i = 0; i < 4; ++i) y[i] = st[i] + p.dt2 * k1[i];
  acrobot_dsdt(&p, y, torque, k2);
  for (int i = 0; i < 4; ++i) y[i] = st[i] + p.dt2 * k2[i];
  acrobot_dsdt(&p, y, torque, k3);
  for (int i = 0; i < 4; ++i) y[i] = st[i] + p.dt * k3[i];
  acrobot_dsdt(&p, y, torque, k4);
  for (int i = 0; i < 4; ++i) {
    float acc = (k1[i] + 2.0f * k2[i]) + 2.0f * k3[i];
    acc = acc + k4[i];
    y[i] = st[i] + p.dt6 * acc;
  }
  y[0] = wrapf(y[0], -ORACLE_PI_F, ORACLE_PI_F);
  y[1] = wrapf(y[1], -ORACLE_PI_F, ORACLE_PI_F);
  y[2] = boundf(y[2], -p.max_vel_1, p.max_vel_1);
  y[3] = boundf(y[3], -p.max_vel_2, p.max_vel_2);
  for (int i = 0; i < 4; ++i) st[i] = y[i];
  int terminated = (-ref_cosf(y[0]) - ref_cosf(y[1] + y[0])) > 1.0f;
  *reward = terminated ? 0.0f : -1.0f;
  return (terminated ? ORACLE_FLAG_TERMINATED : 0u) | time_limit(cfg, steps);
}

/* ------------------------------------------------------------------------- */
/* kind dispatch                                                              */
/* ------------------------------------------------------------------------- */
int oracle_state_dim(int kind) {
  static const int d[ORACLE_NUM_KINDS] = {4, 2, 2, 2, 4};
  return (kind >= 0 && kind < ORACLE_NUM_KINDS) ? d[kind] : -1;
}
int oracle_obs_dim(int kind) {
  static const int d[ORACLE_NUM_KINDS] = {4, 2, 2, 3, 6};
  return (kind >= 0 && kind < ORACLE_NUM_KINDS) ? d[kind] : -1;
}
int oracle_action_is_continuous(int kind) {
  return kind == ORACLE_MOUNTAIN_CAR_CONTINUOUS_V0 || kind == ORACLE_PENDULUM_V1;
}

void oracle_config_default(int kind, oracle_config *cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->auto_reset = 0;
  cfg->is_euler = 1;            /* cartpole.rs:40 */
  cfg->sutton_barto_reward = 0; /* cartpole.rs:39 */
  cfg->goal_velocity = 0.0f;    /* mountain_car.rs:33 */
  switch (kind) {               /* Gymnasium registrations for the envs the reference lacks */
    case ORACLE_MOUNTAIN_CAR_CONTINUOUS_V0: cfg->max_episode_steps = 999; break;
    case ORACLE_PENDULUM_V1: cfg->max_episode_steps = 200; break;
    case ORACLE_ACROBOT_V1: cfg->max_episode_steps = 500; break;
    default: cfg->max_episode_steps = 0; break; /* CartPole: built in; MountainCar: never truncates */
  }
}

static inline uint32_t env_step(int kind, const oracle_config *cfg, float *st, uint32_t *steps,
                                uint32_t *sbt, uint32_t au, float af, float *reward) {
  switch (kind) {
    case ORACLE_CARTPOLE_V1: return cartpole_step(cfg, st, steps, sbt, au, reward);
    case ORACLE_MOUNTAIN_CAR_V0: return mountain_car_step(cfg, st, steps, au, reward);
    case ORACLE_MOUNTAIN_CAR_CONTINUOUS_V0: return mountain_car_continuous_step(cfg, st, steps, af, reward);
    case ORACLE_PENDULUM_V1: return pendulum_step(cfg, st, steps, af, reward);
    default: return acrobot_step(cfg, st, steps, au, reward);
  }
}

static inline void env_obs(int kind, const float *st, float *obs) {
  switch (kind) {
    case ORACLE_PENDULUM_V1:
      obs[0] = ref_cosf(st[0]);
      obs[1] = ref_sinf(st[0]);
      obs[2] = st[1];
      break;
    case ORACLE_ACROBOT_V1:
      obs[0] = ref_cosf(st[0]);
      obs[1] = ref_sinf(st[0]);
      obs[2] = ref_cosf(st[1]);
      obs[3] = ref_sinf(st[1]);
      obs[4] = st[2];
      obs[5] = st[3];
      break;
    default:
      for (int c = 0; c < oracle_state_dim(kind); ++c) obs[c] = st[c];
  }
}

/* reset distributions: cartpole.rs:240, mountain_car.rs:281-283; Gymnasium for the rest */
static inline void reset_from_words(int kind, const uint32_t w[4], float *st) {
  switch (kind) {
    case ORACLE_CARTPOLE_V1:
      for (int c = 0; c < 4; ++c) st[c] = uniform_f64_to_f32(w[c], -0.05, 0.05);
      break;
    case ORACLE_MOUNTAIN_CAR_V0:
    case ORACLE_MOUNTAIN_CAR_CONTINUOUS_V0:
      st[0] = uniform_f64_to_f32(w[0], -0.6, -0.4);
      st[1] = 0.0f;
      break;
    case ORACLE_PENDULUM_V1:
      st[0] = uniform_f64_to_f32(w[0], -ORACLE_PI_D, ORACLE_PI_D);
      st[1] = uniform_f64_to_f32(w[1], -1.0, 1.0);
      break;
    default:
      for (int c = 0; c < 4; ++c) st[c] = uniform_f64_to_f32(w[c], -0.1, 0.1);
  }
}

void oracle_reset_state(int kind, uint64_t seed, uint64_t g, uint64_t t, uint32_t tag, float *state) {
  uint32_t w[4];
  philox_env(seed, g, t, tag, w);
  reset_from_words(kind, w, state);
}

void oracle_sample_action(int kind, uint64_t seed, uint64_t g, uint64_t t, uint8_t *a_u8, float *a_f32) {
  /* one Philox block serves 4 consecutive envs: block index g >> 2, word g & 3 */
  uint32_t w4[4];
  philox_env(seed, g >> 2, t, 2u, w4);
  const uint32_t w = w4[g & 3];
  switch (kind) {
    case ORACLE_CARTPOLE_V1: *a_u8 = (uint8_t)(((uint64_t)w * 2u) >> 32); break;
    case ORACLE_MOUNTAIN_CAR_V0:
    case ORACLE_ACROBOT_V1: *a_u8 = (uint8_t)(((uint64_t)w * 3u) >> 32); break;
    case ORACLE_MOUNTAIN_CAR_CONTINUOUS_V0: *a_f32 = uniform_f64_to_f32(w, -1.0, 1.0); break;
    default: *a_f32 = uniform_f64_to_f32(w, -2.0, 2.0);
  }
}

ORACLE_HOT uint32_t oracle_env_step(int kind, const oracle_config *cfg, float *state, uint32_t *steps,
                                    uint32_t *sbt, uint32_t action_u, float action_f, float *reward) {
  return env_step(kind, cfg, state, steps, sbt, action_u, action_f, reward);
}

ORACLE_HOT void oracle_env_obs(int kind, const float *state, float *obs) { env_obs(kind, state, obs); }

/* ------------------------------------------------------------------------- */
/* batched drivers (semantics of mgym_step / mgym_rollout / mgym_reset)       */
/* ------------------------------------------------------------------------- */

static inline void vec_step_one(int kind, const oracle_config *cfg, uint64_t i, uint64_t ld, uint64_t t,
                                float *state, uint32_t *steps, uint32_t *sbt, float *ep_return,
                                uint32_t au, float af, const float *reset_pool, uint64_t pool_len,
                                float *obs_out, float *reward_out, uint8_t *flags_out,
                                float *final_obs_out, oracle_stats *stats) {
  const int sd = oracle_state_dim(kind), od = oracle_obs_dim(kind);
  float st[4], obs[6], reward;
  uint32_t my_steps = steps ? steps[i] : 0u;
  uint32_t my_sbt = sbt ? sbt[i] : ORACLE_SBT_NONE;
  for (int c = 0; c < sd; ++c) st[c] = state[(uint64_t)c * ld + i];
  if (cfg->auto_reset) my_sbt = ORACLE_SBT_NONE; /* auto-reset presumes reset() precedes every episode */

  uint32_t flags = env_step(kind, cfg, st, &my_steps, &my_sbt, au, af, &reward);
  float ret = 0.0f;
  if (ep_return) ret = ep_return[i] + reward;

  env_obs(kind, st, obs);
  if (final_obs_out)
    for (int c = 0; c < od; ++c) final_obs_out[(uint64_t)c * ld + i] = obs[c];

  if (cfg->auto_reset && flags) {
    if (stats) {
      stats->episodes += 1;
      stats->terminated += (flags & ORACLE_FLAG_TERMINATED) ? 1 : 0;
      stats->truncated += (flags & ORACLE_FLAG_TRUNCATED) ? 1 : 0;
      stats->length_sum += my_steps;
      stats->return_sum += (double)ret;
    }
    uint64_t g = cfg->env_index_base + i;
    if (reset_pool && pool_len) {
      uint64_t j = (g + t) % pool_len;
      for (int c = 0; c < sd; ++c) st[c] = reset_pool[(uint64_t)c * pool_len + j];
    } else {
      /* keyed by the step index at which the finished episode BEGAN (t + 1 - its length): known for the whole
         episode, so the device may draw the state ahead of time (rollout_kernel); see mgym.h "Reset states" */
      oracle_reset_state(kind, cfg->seed, g, t + 1u - (uint64_t)my_steps, 0u, st);
    }
    my_steps = 0;
    my_sbt = ORACLE_SBT_NONE;
    ret = 0.0f;
    env_obs(kind, st, obs);
  }
  for (int c = 0; c < sd; ++c) state[(uint64_t)c * ld + i] = st[c];
  if (steps) steps[i] = my_steps;
  if (sbt) sbt[i] = my_sbt;
  if (ep_return) ep_return[i] = ret;
  if (obs_out)
    for (int c = 0; c < od; ++c) obs_out[(uint64_t)c * ld + i] = obs[c];
  if (reward_out) reward_out[i] = reward;
  if (flags_out) flags_out[i] = (uint8_t)flags;
}

ORACLE_HOT void oracle_vec_step(int kind, const oracle_config *cfg, uint64_t n, uint64_t ld, uint64_t t,
                                float *state, uint32_t *steps, uint32_t *sbt, float *ep_return,
                                const void *actions, const float *reset_pool, uint64_t pool_len,
                                float *obs_out, float *reward_out, uint8_t *flags_out,
                                float *final_obs_out, oracle_stats *stats) {
  const int cont = oracle_action_is_continuous(kind);
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t au = cont ? 0u : ((const uint8_t *)actions)[i];
    float af = cont ? ((const float *)actions)[i] : 0.0f;
    vec_step_one(kind, cfg, i, ld, t, state, steps, sbt, ep_return, au, af, reset_pool, pool_len,
                 obs_out, reward_out, flags_out, final_obs_out, stats);
  }
}

ORACLE_HOT void oracle_vec_rollout(int kind, const oracle_config *cfg, uint64_t n, uint64_t ld,
                                   uint64_t t0, uint32_t K, float *state, uint32_t *steps,
                                   uint32_t *sbt, float *ep_return, const void *actions,
                                   const float *reset_pool, uint64_t pool_len, float *obs_traj,
                                   float *reward_traj, uint8_t *flags_traj, uint64_t *done_count,
                                   oracle_stats *stats) {
  const int cont = oracle_action_is_continuous(kind), od = oracle_obs_dim(kind);
  uint64_t dones = 0;
  for (uint32_t k = 0; k < K; ++k) {
    float *obs_k = obs_traj ? obs_traj + (uint64_t)k * od * ld : NULL;
    float *rew_k = reward_traj ? reward_traj + (uint64_t)k * ld : NULL;
    uint8_t *flg_k = flags_traj ? flags_traj + (uint64_t)k * ld : NULL;
    for (uint64_t i = 0; i < n; ++i) {
      uint8_t a8 = 0;
      float af = 0.0f;
      if (actions) {
        if (cont) af = ((const float *)actions)[(uint64_t)k * ld + i];
        else a8 = ((const uint8_t *)actions)[(uint64_t)k * ld + i];
      } else {
        oracle_sample_action(kind, cfg->seed, cfg->env_index_base + i, t0 + k, &a8, &af);
      }
      uint8_t f = 0;
      vec_step_one(kind, cfg, i, ld, t0 + k, state, steps, sbt, ep_return, a8, af, reset_pool,
                   pool_len, obs_k, rew_k, &f, NULL, stats);
      if (flg_k) flg_k[i] = f;
      dones += f ? 1 : 0;
    }
  }
  if (done_count) *done_count = dones;
}

ORACLE_HOT void oracle_vec_reset(int kind, const oracle_config *cfg, uint64_t n, uint64_t ld,
                                 uint64_t reset_index, const uint8_t *mask, float *state,
                                 uint32_t *steps, uint32_t *sbt, float *ep_return,
                                 const float *reset_pool, uint64_t pool_len, float *obs_out) {
  const int sd = oracle_state_dim(kind), od = oracle_obs_dim(kind);
  for (uint64_t i = 0; i < n; ++i) {
    float st[4], obs[6];
    if (mask && !mask[i]) {
      if (obs_out) {
        for (int c = 0; c < sd; ++c) st[c] = state[(uint64_t)c * ld + i];
        env_obs(kind, st, obs);
        for (int c = 0; c < od; ++c) obs_out[(uint64_t)c * ld + i] = obs[c];
      }
      continue;
    }
    uint64_t g = cfg->env_index_base + i;
    if (reset_pool && pool_len) {
      uint64_t j = (g + reset_index) % pool_len;
      for (int c = 0; c < sd; ++c) st[c] = reset_pool[(uint64_t)c * pool_len + j];
    } else {
      oracle_reset_state(kind, cfg->seed, g, reset_index, 1u, st);
    }
    for (int c = 0; c < sd; ++c) state[(uint64_t)c * ld + i] = st[c];
    if (steps) steps[i] = 0;                /* cartpole.rs:243 */
    if (sbt) sbt[i] = ORACLE_SBT_NONE;      /* cartpole.rs:239 */
    if (ep_return) ep_return[i] = 0.0f;
    if (obs_out) {
      env_obs(kind, st, obs);
      for (int c = 0; c < od; ++c) obs_out[(uint64_t)c * ld + i] = obs[c];
    }
  }
}

/* ------------------------------------------------------------------------- */
/* CPU baseline: the reference's caller loop, one env per thread              */
/* (cartpole.rs:460-471: sample action, step, reset on done).                 */
/* ------------------------------------------------------------------------- */
typedef struct {
  int kind;
  oracle_config cfg;
  uint64_t steps, thread_index;
  double seconds, checksum;
} baseline_job;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

ORACLE_HOT static void baseline_run(baseline_job *job) {
  const int kind = job->kind;
  const int cont = oracle_action_is_continuous(kind);
  const uint32_t n_act = (kind == ORACLE_CARTPOLE_V1) ? 2u : 3u;
  float st[4] = {0, 0, 0, 0};
  uint32_t steps = 0, sbt = ORACLE_SBT_NONE;
  uint64_t episode = 0, g = job->cfg.env_index_base + job->thread_index;
  uint64_t rng = 0x9E3779B97F4A7C15ull * (g + 1) + job->cfg.seed;
  double acc = 0.0;
  oracle_reset_state(kind, job->cfg.seed, g, episode++, 1u, st);
  double t0 = now_s();
  for (uint64_t s = 0; s < job->steps; ++s) {
    rng ^= rng << 13; /* xorshift64: stands in for action_space.sample() */
    rng ^= rng >> 7;
    rng ^= rng << 17;
    uint32_t w = (uint32_t)(rng >> 32);
    uint32_t au = (uint32_t)(((uint64_t)w * n_act) >> 32);
    float af = cont ? ((float)(int32_t)w * 0x1p-31f) * (kind == ORACLE_PENDULUM_V1 ? 2.0f : 1.0f) : 0.0f;
    float reward;
    uint32_t flags = env_step(kind, &job->cfg, st, &steps, &sbt, au, af, &reward);
    acc += (double)reward + (double)st[0];
    if (flags) {
      oracle_reset_state(kind, job->cfg.seed, g, episode++, 1u, st);
      steps = 0;
      sbt = ORACLE_SBT_NONE;
    }
  }
  job->seconds = now_s() - t0;
  job->checksum = acc;
}

static void *baseline_thread(void *arg) {
  baseline_run((baseline_job *)arg);
  return NULL;
}

uint64_t oracle_baseline_loop(int kind, const oracle_config *cfg, uint64_t steps_per_thread,
                              int n_threads, double *seconds, double *checksum) {
  if (n_threads < 1) n_threads = 1;
  baseline_job *jobs = (baseline_job *)calloc((size_t)n_threads, sizeof(baseline_job));
  pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
  double t0 = now_s();
  for (int i = 0; i < n_threads; ++i) {
    jobs[i].kind = kind;
    jobs[i].cfg = *cfg;
    jobs[i].steps = steps_per_thread;
    jobs[i].thread_index = (uint64_t)i;
    pthread_create(&th[i], NULL, baseline_thread, &jobs[i]);
  }
  double cs = 0.0;
  for (int i = 0; i < n_threads; ++i) {
    pthread_join(th[i], NULL);
    cs += jobs[i].checksum;
  }
  double wall = now_s() - t0;
  if (seconds) *seconds = wall;
  if (checksum) *checksum = cs;
  free(jobs);
  free(th);
  return steps_per_thread * (uint64_t)n_threads;
}
