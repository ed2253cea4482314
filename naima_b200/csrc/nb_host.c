/* nb_host.c -- host-side helper of the device-resident sampler: the random draws of a
 * block of ensemble steps, bit-identical to what emcee's stretch move consumes from a
 * numpy.random.RandomState (legacy MT19937 stream), without the per-call NumPy overhead
 * (12 small-array calls, ~50 us per 256-walker step).
 *
 * Restated from the published algorithms the stream is defined by (the reference reaches
 * them through emcee -> numpy.random.mtrand; nothing of this is in /root/reference):
 *   MT19937 (Matsumoto & Nishimura 1998): state key[624], position pos;
 *   random_sample / rand : (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53 from two 32-bit draws;
 *   shuffle              : Fisher-Yates from the top, j = masked rejection draw in [0, i];
 *   randint(0, n)        : masked rejection on 32-bit draws.
 * Per ensemble step, in emcee's order: one uniform (choice of the move), shuffle of
 * arange(W) % 2, then per half: rand(Ns) -> zz, randint(Ns) -> partner, rand(Ns) -> accept.
 * The caller takes log() of the accept uniforms with NumPy (its log is not libm's).
 * Compile with -ffp-contract=off: zz must round like NumPy's separate multiply and add.
 */
#include <stdint.h>
#include <stdlib.h>

#define MT_N 624
#define MT_M 397

typedef struct {
  uint32_t* key;
  int pos;
} mt_state;

static void mt_gen(mt_state* s) {
  uint32_t* k = s->key;
  uint32_t y;
  int i;
  for (i = 0; i < MT_N - MT_M; i++) {
    y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu);
    k[i] = k[i + MT_M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
  }
  for (; i < MT_N - 1; i++) {
    y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu);
    k[i] = k[i + (MT_M - MT_N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
  }
  y = (k[MT_N - 1] & 0x80000000u) | (k[0] & 0x7fffffffu);
  k[MT_N - 1] = k[MT_M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
  s->pos = 0;
}

static inline uint32_t mt_next(mt_state* s) {
  uint32_t y;
  if (s->pos == MT_N) mt_gen(s);
  y = s->key[s->pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

static inline double mt_double(mt_state* s) {
  int32_t a = (int32_t)(mt_next(s) >> 5), b = (int32_t)(mt_next(s) >> 6);
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

static inline uint32_t mask_of(uint32_t max) {
  uint32_t m = max;
  m |= m >> 1;
  m |= m >> 2;
  m |= m >> 4;
  m |= m >> 8;
  m |= m >> 16;
  return m;
}

/* uniform integer in [0, max] by masked rejection on 32-bit draws */
static inline uint32_t mt_interval(mt_state* s, uint32_t max) {
  uint32_t mask, v;
  if (max == 0) return 0;
  mask = mask_of(max);
  while ((v = (mt_next(s) & mask)) > max) {
  }
  return v;
}

/* key[624], *pos: the RandomState's MT19937 state, advanced in place.
 * Outputs [nsteps][2][Ns] with Ns = W / 2 (W even): s_idx, c_idx (int32), zz, u_acc.
 * Returns 0, or -1 on bad arguments / allocation failure. */
int nb_host_draw_steps(uint32_t* key, int* pos, int W, int nsteps, double a, int32_t* s_idx,
                       int32_t* c_idx, double* zz, double* u_acc) {
  mt_state s;
  int Ns = W / 2, t, i, split;
  unsigned char* inds;
  int32_t* half[2];
  if (!key || !pos || W < 2 || (W & 1) || nsteps < 0 || *pos < 0 || *pos > MT_N) return -1;
  inds = (unsigned char*)malloc((size_t)W);
  half[0] = (int32_t*)malloc(sizeof(int32_t) * (size_t)W);
  if (!inds || !half[0]) {
    free(inds);
    free(half[0]);
    return -1;
  }
  half[1] = half[0] + Ns;
  s.key = key;
  s.pos = *pos;
  for (t = 0; t < nsteps; ++t) {
    size_t base = (size_t)t * 2 * Ns;
    int n0 = 0, n1 = 0;
    (void)mt_double(&s); /* random.choice(moves, p=weights): one uniform */
    for (i = 0; i < W; ++i) inds[i] = (unsigned char)(i & 1);
    for (i = W - 1; i >= 1; --i) { /* RandomState.shuffle */
      uint32_t j = mt_interval(&s, (uint32_t)i);
      unsigned char tmp = inds[i];
      inds[i] = inds[j];
      inds[j] = tmp;
    }
    for (i = 0; i < W; ++i) {
      if (inds[i] == 0) {
        if (n0 < Ns) half[0][n0] = i;
        n0++;
      } else {
        if (n1 < Ns) half[1][n1] = i;
        n1++;
      }
    }
    for (split = 0; split < 2; ++split) {
      size_t o = base + (size_t)split * Ns;
      const int32_t* mine = half[split];
      const int32_t* other = half[1 - split];
      for (i = 0; i < Ns; ++i) s_idx[o + i] = mine[i];
      for (i = 0; i < Ns; ++i) {
        double r = mt_double(&s);
        double v = (a - 1.0) * r;
        v = v + 1.0;
        v = v * v;
        zz[o + i] = v / a;
      }
      for (i = 0; i < Ns; ++i) c_idx[o + i] = other[mt_interval(&s, (uint32_t)(Ns - 1))];
      for (i = 0; i < Ns; ++i) u_acc[o + i] = mt_double(&s);
    }
  }
  *pos = s.pos;
  free(inds);
  free(half[0]);
  return 0;
}
