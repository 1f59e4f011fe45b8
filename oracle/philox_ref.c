/* TEST INFRASTRUCTURE ONLY -- plain C restatement of the engine's Philox4x32-10
 * index sampler (jsrl_corl_b200/csrc/common.cuh) and of the reference's
 * ReplayBuffer.sample gather (algorithms/finetune/iql.py:171-178) on packed
 * rows.  Compiled by __graft_entry__.build() into oracle/_build/.  Only tests,
 * smoke() and bench.py's cpu_baseline may load it. */
#include <stdint.h>
#include <string.h>

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

void philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, key[0], key[1]);
  memcpy(out, c, sizeof(c));
}

void philox_indices(uint64_t seed, uint64_t step, uint64_t size, int64_t batch, int64_t* out) {
  for (int64_t b = 0; b < batch; ++b) {
    uint32_t c[4] = {(uint32_t)(b >> 1), (uint32_t)step, (uint32_t)(step >> 32), 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t u = (b & 1) ? (((uint64_t)c[3] << 32) | c[2]) : (((uint64_t)c[1] << 32) | c[0]);
    out[b] = (int64_t)(((unsigned __int128)u * (unsigned __int128)size) >> 64);
  }
}

/* gather `batch` packed rows (row_floats each) by index: the CPU statement of the replay gather */
void gather_rows(const float* rows, int64_t row_floats, const int64_t* idx, int64_t batch, float* out) {
  for (int64_t b = 0; b < batch; ++b) memcpy(out + b * row_floats, rows + idx[b] * row_floats, sizeof(float) * row_floats);
}
