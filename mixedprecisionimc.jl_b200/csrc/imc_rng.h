// imc_rng.h — counter-based random draws for host and device.
//
// The reference draws from Julia's global task-local RNG (Random.seed!, MixedPrecisionIMC.jl:86),
// which makes every history depend on loop order.  The engine replaces it with Philox4x32-10
// (Salmon et al., SC'11) keyed by the deck SEED and counted by (particle id, time step, stream,
// block), so a particle's draws do not depend on which GPU or thread tracks it (BASELINE.json
// north_star "RNG").  Replay mode bypasses Philox and reads pre-drawn numbers from a tape.
//
// Draw conversions follow Julia's conventions: rand(T) is a multiple of 2^-11 / 2^-24 / 2^-53
// in [0,1); randexp(T) for T < Float64 is drawn wider and converted (here: -log of a 32-bit
// uniform in Float32, by the division-free dm::neglog_unit_f; Julia draws Float64 — statistically equivalent, not bit-equivalent; the
// Julia stream itself is not reproducible across Julia versions).
#pragma once
#include "imc_num.h"
#include "imc_math.h"

namespace imc {

struct Philox {
  static IMC_HD void mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *lo = (uint32_t)p;
    *hi = (uint32_t)(p >> 32);
#endif
  }
  // Philox4x32-10 block function
  static IMC_HD void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, &hi0, &lo0);
      mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
      uint32_t n0 = hi1 ^ c1 ^ k0;
      uint32_t n2 = hi0 ^ c3 ^ k1;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
  // the same block function with the ten round keys precomputed (kernel-uniform: they sit in the constant bank
  // and enter the xor as an operand, instead of two additions per round per thread)
  static IMC_HD void round_keys(uint64_t seed, uint32_t rk[20]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  }
  static IMC_HD void block_rk(const uint32_t ctr[4], const uint32_t* rk, uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, &hi0, &lo0);
      mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
      uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
      uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

enum : uint32_t { STREAM_SOURCE = 0u, STREAM_TRACK = 1u, STREAM_TRACK_EXTRA = 2u, STREAM_PLANCK = 3u };

// One particle's draw stream for one time step.  Words are consumed in order; a new Philox
// block is generated every 4 words.  counter = (id_lo, id_hi, stream<<28 | step, block): the block index has a whole
// word to itself (2^32 blocks per history and stream), the time step 28 bits (the engines refuse step >= 2^28).
struct PhiloxStream {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t buf[4];
  uint32_t used;  // words consumed from buf (4 = empty)
  IMC_HD void init(uint64_t seed, uint64_t id, uint32_t step, uint32_t stream) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    ctr[0] = (uint32_t)id;
    ctr[1] = (uint32_t)(id >> 32);
    ctr[2] = step | (stream << 28);
    ctr[3] = 0u;
    buf[0] = buf[1] = buf[2] = buf[3] = 0u;
    used = 4;
  }
  IMC_HD uint32_t next_u32() {
    if (used == 4) {
      Philox::block(ctr, key, buf);
      ctr[3] += 1;
      used = 0;
    }
    // avoid dynamic register-array indexing
    uint32_t w = used == 0 ? buf[0] : used == 1 ? buf[1] : used == 2 ? buf[2] : buf[3];
    used += 1;
    return w;
  }
  IMC_HD uint64_t next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return (hi << 32) | lo;
  }
};

// word -> uniform / exponential conversions (shared by every RNG back-end)
template <class P> IMC_HD Num<P> uniform_from_word(uint64_t w);
template <> IMC_HD Num<F16> uniform_from_word<F16>(uint64_t w) { return Num<F16>((float)((uint32_t)w >> 21) * 4.8828125e-04f); }
template <> IMC_HD Num<F32> uniform_from_word<F32>(uint64_t w) { return Num<F32>((float)((uint32_t)w >> 8) * 5.9604644775390625e-08f); }
template <> IMC_HD Num<F64> uniform_from_word<F64>(uint64_t w) { return Num<F64>((double)(w >> 11) * 1.1102230246251565e-16); }

IMC_HD double randexp64_from_word(uint64_t w) {
  double u = ((double)(w >> 11) + 1.0) * 1.1102230246251565e-16;  // (0, 1]
  return -dm::log_d(u);
}
IMC_HD float randexp32_from_word(uint32_t w) {
  float u = (float)w * 2.3283064365386963e-10f + 1.1641532182693481e-10f;  // (0, 1]
  if (u > 1.0f) u = 1.0f;
  return dm::neglog_unit_f(u);  // u in [2^-33, 1]: positive and normal
}

// RNG back-end 1: Philox.  Draw<P> API: uniform() -> rand(T); randexp() -> randexp(T);
// randexp64() -> randexp() in Float64 (MC_RW, imc_transport.jl:279).
template <class P>
struct PhiloxDraw {
  PhiloxStream s;
  IMC_HD void init(uint64_t seed, uint64_t id, uint32_t step, uint32_t stream) { s.init(seed, id, step, stream); }
  IMC_HD Num<P> uniform() {
    if constexpr (P::id == 2) return uniform_from_word<P>(s.next_u64());
    else return uniform_from_word<P>((uint64_t)s.next_u32());
  }
  IMC_HD Num<P> randexp() {
    if constexpr (P::id == 2) return Num<P>(randexp64_from_word(s.next_u64()));
    else return Num<P>(P::rnd(randexp32_from_word(s.next_u32())));
  }
  IMC_HD double randexp64() { return randexp64_from_word(s.next_u64()); }
  IMC_HD bool exhausted() const { return false; }
};

// RNG back-end 1b: Philox with draws RESERVED PER SEGMENT, for the history loops of MC / MC2D.
// Every loop iteration draws exactly one exponential (imc_transport.jl:87, :561) and at most one uniform (the
// new direction after a collision, :180, :708), so segment n owns fixed words of the particle's stream:
//   Float16/Float32: block n>>1 = [exp(n), uni(n), exp(n+1), uni(n+1)]      (one Philox block per two segments)
//   Float64        : block n    = [exp lo, exp hi, uni lo, uni hi]           (one block per segment)
// All lanes of a warp therefore generate their blocks in the same iterations (no divergent refresh) and a
// particle's draws still depend only on (seed, particle id, step, segment).  The rare extra draws of the 1-D
// `while mu == 0` resampling come from a separate sequential stream.
template <class P>
struct SegDraw {
  uint32_t buf[4];
  uint32_t id_lo, id_hi;
  uint32_t n;        // current segment (0-based); 0xffffffff before the first
  uint32_t extra_n;  // bit 31: the segment's reserved uniform is used; low bits: extra words consumed so far
  IMC_HD void init(uint64_t id) {
    id_lo = (uint32_t)id; id_hi = (uint32_t)(id >> 32);
    buf[0] = buf[1] = buf[2] = buf[3] = 0u;
    n = 0xffffffffu;
    extra_n = 0u;
  }
  // seed / step are passed in (kernel-uniform values) instead of being carried per thread
  // continue a history at segment `next` (event-based tracking reloads the particle every segment): the block
  // of the current pair is regenerated on the next call whatever its parity
  IMC_HD void resume(uint64_t id, uint32_t next, uint32_t extra_words) {
    init(id);
    n = next - 1u;
    extra_n = extra_words | 0x40000000u;  // bit 30: buffer invalid
  }
  IMC_HD void next_segment(uint64_t seed, uint32_t step) {
    n += 1u;
    const bool stale = (extra_n & 0x40000000u) != 0u;
    extra_n &= 0x3fffffffu;
    if (P::id == 2 || (n & 1u) == 0u || stale) {
      uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
      uint32_t c[4] = {id_lo, id_hi, step | (STREAM_TRACK << 28), P::id == 2 ? n : (n >> 1)};
      Philox::block(c, key, buf);
    }
  }
  IMC_HD void next_segment_rk(const uint32_t* rk, uint32_t step) {  // as next_segment, round keys precomputed
    n += 1u;
    const bool stale = (extra_n & 0x40000000u) != 0u;
    extra_n &= 0x3fffffffu;
    if (P::id == 2 || (n & 1u) == 0u || stale) {
      uint32_t c[4] = {id_lo, id_hi, step | (STREAM_TRACK << 28), P::id == 2 ? n : (n >> 1)};
      Philox::block_rk(c, rk, buf);
    }
  }
  IMC_HD Num<P> randexp() const {
    if constexpr (P::id == 2) return Num<P>(randexp64_from_word(((uint64_t)buf[1] << 32) | buf[0]));
    else return Num<P>(P::rnd(randexp32_from_word((n & 1u) ? buf[2] : buf[0])));
  }
  IMC_HD uint32_t extra_word(uint64_t seed, uint32_t step) {  // sequential words of the extra stream, block made on demand
    uint32_t j = extra_n & 0x3fffffffu;
    extra_n += 1u;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c[4] = {id_lo, id_hi, step | (STREAM_TRACK_EXTRA << 28), j >> 2}, o[4];
    Philox::block(c, key, o);
    return (j & 3u) == 0u ? o[0] : (j & 3u) == 1u ? o[1] : (j & 3u) == 2u ? o[2] : o[3];
  }
  IMC_HD Num<P> uniform(uint64_t seed, uint32_t step) {
    if (extra_n & 0x80000000u) {  // second and later uniforms of one segment (1-D `while mu == 0`)
      if constexpr (P::id == 2) { uint64_t lo = extra_word(seed, step); uint64_t hi = extra_word(seed, step); return uniform_from_word<P>((hi << 32) | lo); }
      else return uniform_from_word<P>((uint64_t)extra_word(seed, step));
    }
    extra_n |= 0x80000000u;
    if constexpr (P::id == 2) return uniform_from_word<P>(((uint64_t)buf[3] << 32) | buf[2]);
    else return uniform_from_word<P>((uint64_t)((n & 1u) ? buf[3] : buf[1]));
  }
};

// RNG back-end 1c: SegDraw without cursor state, for MC2D — a 2-D segment draws exactly one exponential and at most
// one uniform (no `while mu == 0` resampling), so the caller's segment counter selects the words and there is no
// extra stream.  Same words as SegDraw for the same (seed, particle id, step, segment).
template <class P>
struct SegDrawLean {
  uint32_t buf[4];
  uint32_t id_lo, id_hi;
  IMC_HD void init(uint64_t id) {
    id_lo = (uint32_t)id; id_hi = (uint32_t)(id >> 32);
    buf[0] = buf[1] = buf[2] = buf[3] = 0u;
  }
  // PAR: the parity of `seg` when the caller knows it at compile time (loops unrolled by two segments), else -1
  template <int PAR = -1>
  IMC_HD void next_segment_rk(const uint32_t* rk, uint32_t step, uint32_t seg) {   // seg: 0-based segment of the history
    if (P::id == 2 || (PAR < 0 ? (seg & 1u) == 0u : PAR == 0)) {
      uint32_t c[4] = {id_lo, id_hi, step | (STREAM_TRACK << 28), P::id == 2 ? seg : (seg >> 1)};
      Philox::block_rk(c, rk, buf);
    }
  }
  template <int PAR = -1>
  IMC_HD Num<P> randexp(uint32_t seg) const {
    if constexpr (P::id == 2) return Num<P>(randexp64_from_word(((uint64_t)buf[1] << 32) | buf[0]));
    else return Num<P>(P::rnd(randexp32_from_word((PAR < 0 ? (seg & 1u) != 0u : PAR == 1) ? buf[2] : buf[0])));
  }
  template <int PAR = -1>
  IMC_HD Num<P> uniform(uint32_t seg) const {
    if constexpr (P::id == 2) return uniform_from_word<P>(((uint64_t)buf[3] << 32) | buf[2]);
    else return uniform_from_word<P>((uint64_t)((PAR < 0 ? (seg & 1u) != 0u : PAR == 1) ? buf[3] : buf[1]));
  }
};

// RNG back-end 2: tape (replay mode).  Pre-drawn Float64 numbers, draw-major layout
// tape[k * stride + slot]; uniforms must already be T-representable (they are rand(T) values),
// exponentials are Float64 and are converted to T here exactly as randexp(T) does.
template <class P>
struct TapeDraw {
  const double* uni; const double* ex;
  size_t stride, slot;
  int n_uni, n_exp, iu, ie;
  bool over;
  IMC_HD void init(const double* uni_, int n_uni_, const double* ex_, int n_exp_, size_t stride_, size_t slot_) {
    uni = uni_; ex = ex_; n_uni = n_uni_; n_exp = n_exp_; stride = stride_; slot = slot_;
    iu = 0; ie = 0; over = false;
  }
  IMC_HD Num<P> uniform() {
    if (iu >= n_uni) { over = true; return Num<P>::from_d(0.5); }
    return Num<P>::from_d(uni[(size_t)(iu++) * stride + slot]);
  }
  IMC_HD double randexp64() {
    if (ie >= n_exp) { over = true; return 1.0; }
    return ex[(size_t)(ie++) * stride + slot];
  }
  IMC_HD Num<P> randexp() { return Num<P>::from_d(randexp64()); }
  IMC_HD bool exhausted() const { return over; }
};

}  // namespace imc
