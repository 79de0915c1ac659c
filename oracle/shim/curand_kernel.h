#pragma once
struct curandState { unsigned long long s; };
inline void curand_init(unsigned long long seed, unsigned long long seq, unsigned long long, curandState *st) { st->s = seed * 6364136223846793005ULL + seq * 1442695040888963407ULL + 1; }
inline float curand_uniform(curandState *st) {
  st->s = st->s * 6364136223846793005ULL + 1442695040888963407ULL;
  return ((st->s >> 40) + 1) * (1.0f / 16777216.0f);
}
