// Entropy coding of the latents: quantised-CDF tables, symbol / index preparation and a chunked
// rANS coder, one independent stream per chunk and GPU thread.
//
// Replaces what the reference reaches through its dependency compressai==1.2.4 (not vendored; algorithm restated in
// oracle/rans.py) from task-oriented-PTQ/models/nic_cvt.py:426-570 and light-uniform-PTQ/models/tinylic.py:236-367:
//   pmf_to_quantized_cdf                     (cpp_exts/ops/ops.cpp)               -> b200lic_pmf_to_quantized_cdf (host)
//   EntropyModel.quantize(.., "symbols") + GaussianConditional.build_indexes        -> b200lic_rans_symbols
//   BufferedRansEncoder.encode_with_indexes  (cpp_exts/rans/rans_interface.cpp)     -> b200lic_rans_encode_sizes / _write
//   RansDecoder.decode_with_indexes                                                 -> b200lic_rans_decode
// The coder is ryg_rans rans64 as the interface drives it: 64-bit state, 32-bit renormalisation words, 16-bit
// probabilities, symbols outside a table's support escaped through its last entry and sent as 4-bit bypass digits.  A
// rANS stream is one dependency chain, so the symbol sequence is cut into fixed-size chunks, each coded as a complete
// stream of that format (state flush included) by one thread; an offset table in front of the payload lets the decoder
// start every chunk at once.  With one chunk the payload is the sequential stream itself.  The cost of the cut is one
// flushed state (8 bytes) + one table entry (4 bytes) per chunk.
#include <math.h>
#include <stdint.h>
#include <vector>
#include "common.cuh"

namespace b200lic {

constexpr int kPrecision = 16, kBypassBits = 4, kMaxBypass = 15;
constexpr uint64_t kRansL = 1ull << 31;

__global__ void __launch_bounds__(256)
    rans_symbols_kernel(const float* __restrict__ x, const float* __restrict__ means, int means_per_channel,
                        const float* __restrict__ scales, const float* __restrict__ table, int levels, float bound, int C,
                        int HW, size_t n, int* __restrict__ symbols, int* __restrict__ indexes) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / (size_t)HW) % (size_t)C);
    float v = x[i];
    if (means) v = __fsub_rn(v, means_per_channel ? means[c] : means[i]);
    symbols[i] = (int)rintf(v);                                    // torch.round: half to even
    int idx = c;
    if (scales) {
      // build_indexes: levels - 1 - #{t in table[:-1] : scale <= t}; the table ascends, so the count is a suffix length
      const float s = fmaxf(scales[i], bound);
      int lo = 0, hi = levels - 1;                                 // first j in [0, levels-1) with s <= table[j]
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s <= table[mid]) hi = mid; else lo = mid + 1;
      }
      idx = lo;                                                    // = levels - 1 - (levels - 1 - lo)
    }
    indexes[i] = idx;
  }
}

struct RansEnc {
  uint64_t x;
  unsigned* ptr;        // next word is written at *--ptr (streams are produced back to front, like ryg's encoder)
  unsigned count;
  template <bool WRITE>
  __device__ __forceinline__ void emit() {
    if (WRITE) *--ptr = (unsigned)x;
    ++count;
    x >>= 32;
  }
  // Rans64EncPut, scale_bits = 16.  x / freq is the chain every symbol waits for, and a 64-bit integer division is a
  // ~100-instruction subroutine here; `rd` = 1.0 / freq is computed off the chain (eight symbols at a time), the quotient
  // estimate floor(x * rd) is within one of the true quotient (x < 2^47 * freq after the renormalisation, so the estimate
  // errs by < 2^-4), and the remainder settles it exactly.
  template <bool WRITE>
  __device__ __forceinline__ void put(unsigned start, unsigned freq, double rd) {
    if (x >= ((uint64_t)freq << 47)) emit<WRITE>();
    uint64_t q = __double2ull_rz(__ull2double_rz(x) * rd);
    long long r = (long long)(x - q * freq);
    if (r < 0) {
      --q;
      r += freq;
    } else if (r >= (long long)freq) {
      ++q;
      r -= freq;
    }
    x = (q << kPrecision) + (uint64_t)r + start;
  }
  template <bool WRITE>
  __device__ __forceinline__ void put_bits(unsigned val) {                      // Rans64EncPutBits, 4 bits
    if (x >= ((kRansL >> kBypassBits) << 32)) emit<WRITE>();
    x = (x << kBypassBits) | val;
  }
};

template <bool WRITE>
__global__ void __launch_bounds__(32)
    rans_encode_kernel(const int* __restrict__ sym, const int* __restrict__ idx, unsigned n, unsigned chunk,
                       unsigned n_chunks, const int* __restrict__ cdf, const int* __restrict__ cdf_len,
                       const int* __restrict__ offset, int stride, const unsigned* __restrict__ chunk_off,
                       unsigned* __restrict__ chunk_words, unsigned* __restrict__ out) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const unsigned lo = c * chunk, hi = min(n, lo + chunk);
  RansEnc e;
  e.x = kRansL;
  e.ptr = WRITE ? out + chunk_off[c + 1] : nullptr;
  e.count = 0;
  // Only the state update depends on the previous symbol; every table look-up depends on (symbol, index) alone.  Eight
  // symbols are resolved to (start, freq, escape) with their loads in flight together, then pushed through the state one
  // after the other: the chain per symbol is the 64-bit division, not four dependent global loads.
  constexpr int kBatch = 8;
  for (unsigned top = hi; top > lo;) {
    const unsigned cnt = min((unsigned)kBatch, top - lo);
    int k[kBatch], value[kBatch], max_value[kBatch];
    unsigned raw[kBatch], start[kBatch], freq[kBatch];
    double rd[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const unsigned i = top - 1 - (j < (int)cnt ? j : 0);
      k[j] = idx[i];
      value[j] = sym[i];
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      max_value[j] = cdf_len[k[j]] - 2;
      value[j] -= offset[k[j]];
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      raw[j] = 0;
      if (value[j] < 0) {
        raw[j] = (unsigned)(-2 * (long long)value[j] - 1);
        value[j] = max_value[j];
      } else if (value[j] >= max_value[j]) {
        raw[j] = 2u * (unsigned)(value[j] - max_value[j]);
        value[j] = max_value[j];
      }
      const int* row = cdf + (size_t)k[j] * stride;
      start[j] = (unsigned)row[value[j]];
      freq[j] = (unsigned)row[value[j] + 1] - start[j];
      rd[j] = 1.0 / (double)freq[j];
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      if (j < (int)cnt) {
        if (value[j] == max_value[j]) {
          // forward order: symbol, digit count, digits low to high; coded back to front.  raw has 32 bits: at most 8
          // digits, so the count is a single prefix digit (< 15)
          int n_bypass = 0;
          while (n_bypass < 8 && (raw[j] >> (n_bypass * kBypassBits)) != 0) ++n_bypass;
          for (int d = n_bypass - 1; d >= 0; --d) e.put_bits<WRITE>((raw[j] >> (d * kBypassBits)) & kMaxBypass);
          e.put_bits<WRITE>((unsigned)n_bypass);
        }
        e.put<WRITE>(start[j], freq[j], rd[j]);
      }
    }
    top -= cnt;
  }
  // Rans64EncFlush: low word first in memory
  if (WRITE) {
    *--e.ptr = (unsigned)(e.x >> 32);
    *--e.ptr = (unsigned)e.x;
  } else {
    chunk_words[c] = e.count + 2;
  }
}

__global__ void __launch_bounds__(32)
    rans_decode_kernel(const unsigned* __restrict__ words, const unsigned* __restrict__ chunk_off, unsigned n,
                       unsigned chunk, unsigned n_chunks, const int* __restrict__ idx, const int* __restrict__ cdf,
                       const int* __restrict__ cdf_len, const int* __restrict__ offset, int stride,
                       int* __restrict__ symbols) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const unsigned lo = c * chunk, hi = min(n, lo + chunk);
  const unsigned* ptr = words + chunk_off[c];
  uint64_t x = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
  ptr += 2;
  auto get_bits = [&]() -> unsigned {
    const unsigned v = (unsigned)x & kMaxBypass;
    x >>= kBypassBits;
    if (x < kRansL) x = (x << 32) | *ptr++;
    return v;
  };
  for (unsigned i = lo; i < hi; ++i) {
    const int k = idx[i];
    const int* row = cdf + (size_t)k * stride;
    const int len = cdf_len[k], max_value = len - 2;
    const int cum = (int)((unsigned)x & 0xffffu);
    int a = 0, b = len - 1;                                        // first entry > cum lies in (a, b]: row[0] = 0 <= cum
    while (b - a > 1) {
      const int mid = (a + b) >> 1;
      if (row[mid] > cum) b = mid; else a = mid;
    }
    const unsigned start = (unsigned)row[a], freq = (unsigned)row[a + 1] - start;
    x = (uint64_t)freq * (x >> kPrecision) + ((unsigned)x & 0xffffu) - start;
    if (x < kRansL) x = (x << 32) | *ptr++;
    int value = a;
    if (value == max_value) {
      unsigned val = get_bits();
      unsigned n_bypass = val;
      while (val == (unsigned)kMaxBypass) {
        val = get_bits();
        n_bypass += val;
      }
      unsigned raw = 0;
      for (unsigned j = 0; j < n_bypass; ++j) raw |= get_bits() << (j * kBypassBits);
      value = (int)(raw >> 1);
      if (raw & 1u) value = -value - 1; else value += max_value;
    }
    symbols[i] = value + offset[k];
  }
}

// ops.cpp pmf_to_quantized_cdf for one row (host): pmf[0..n) -> cdf[0..n]
static bool quantized_cdf_row(const float* pmf, int n, int* cdf) {
  std::vector<uint32_t> c((size_t)n + 1);
  c[0] = 0;
  uint32_t total = 0;
  for (int i = 0; i < n; ++i) {
    c[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << kPrecision));
    total += c[i + 1];
  }
  if (total == 0) return false;
  for (auto& v : c) v = (uint32_t)((((uint64_t)1 << kPrecision) * v) / total);
  for (int i = 1; i <= n; ++i) c[i] += c[i - 1];
  c[n] = 1u << kPrecision;
  for (int i = 0; i < n; ++i) {
    if (c[i] != c[i + 1]) continue;
    uint32_t best_freq = ~0u;                                       // steal from the least frequent symbol that can spare one
    int best = -1;
    for (int j = 0; j < n; ++j) {
      const uint32_t f = c[j + 1] - c[j];
      if (f > 1 && f < best_freq) {
        best_freq = f;
        best = j;
      }
    }
    if (best < 0) return false;
    if (best < i) {
      for (int j = best + 1; j <= i; ++j) --c[j];
    } else {
      for (int j = i + 1; j <= best; ++j) ++c[j];
    }
  }
  for (int i = 0; i <= n; ++i) cdf[i] = (int)c[i];
  return true;
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

int b200lic_pmf_to_quantized_cdf(const float* pmf, const float* tail_mass, const int* pmf_length, int rows, int max_length,
                                 int* cdf_out) {
  B200_REQUIRE(pmf && tail_mass && pmf_length && cdf_out && rows > 0 && max_length > 0, "pmf_to_quantized_cdf: bad arguments");
  std::vector<float> prob((size_t)max_length + 1);
  for (int r = 0; r < rows; ++r) {
    const int n = pmf_length[r];
    B200_REQUIRE(n > 0 && n <= max_length, "pmf_to_quantized_cdf: row %d has length %d (max %d)", r, n, max_length);
    for (int i = 0; i < n; ++i) prob[i] = pmf[(size_t)r * max_length + i];
    prob[n] = tail_mass[r];
    int* row = cdf_out + (size_t)r * (max_length + 2);
    for (int i = 0; i < max_length + 2; ++i) row[i] = 0;
    if (!quantized_cdf_row(prob.data(), n + 1, row)) {
      set_error("pmf_to_quantized_cdf: row %d cannot be normalised", r);
      return B200LIC_ERR_ARG;
    }
  }
  return B200LIC_OK;
}

int b200lic_rans_symbols(const float* x, const float* means, int means_per_channel, const float* scales,
                         const float* scale_table, int levels, float scale_bound, int C, int HW, size_t n, int* symbols,
                         int* indexes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && symbols && indexes && C > 0 && HW > 0, "rans_symbols: bad arguments");
  B200_REQUIRE(!scales || (scale_table && levels >= 2), "rans_symbols: scales need a scale table");
  if (n == 0) return B200LIC_OK;
  rans_symbols_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, means, means_per_channel, scales, scale_table,
                                                                     levels, scale_bound, C, HW, n, symbols, indexes);
  B200_LAUNCH_CHECK("rans_symbols_kernel");
  return B200LIC_OK;
}

int b200lic_rans_encode_sizes(const int* symbols, const int* indexes, unsigned n, unsigned chunk, const int* cdf,
                              const int* cdf_len, const int* offset, int cdf_stride, unsigned* chunk_words,
                              b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(symbols && indexes && cdf && cdf_len && offset && chunk_words && chunk > 0 && cdf_stride > 2,
               "rans_encode_sizes: bad arguments");
  if (n == 0) return B200LIC_OK;
  const unsigned n_chunks = (n + chunk - 1) / chunk;
  rans_encode_kernel<false><<<(n_chunks + 31) / 32, 32, 0, as_stream(stream)>>>(symbols, indexes, n, chunk, n_chunks, cdf,
                                                                              cdf_len, offset, cdf_stride, nullptr,
                                                                              chunk_words, nullptr);
  B200_LAUNCH_CHECK("rans_encode_kernel<sizes>");
  return B200LIC_OK;
}

int b200lic_rans_encode_write(const int* symbols, const int* indexes, unsigned n, unsigned chunk, const int* cdf,
                              const int* cdf_len, const int* offset, int cdf_stride, const unsigned* chunk_off,
                              unsigned* out_words, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(symbols && indexes && cdf && cdf_len && offset && chunk_off && out_words && chunk > 0 && cdf_stride > 2,
               "rans_encode_write: bad arguments");
  if (n == 0) return B200LIC_OK;
  const unsigned n_chunks = (n + chunk - 1) / chunk;
  rans_encode_kernel<true><<<(n_chunks + 31) / 32, 32, 0, as_stream(stream)>>>(symbols, indexes, n, chunk, n_chunks, cdf,
                                                                             cdf_len, offset, cdf_stride, chunk_off, nullptr,
                                                                             out_words);
  B200_LAUNCH_CHECK("rans_encode_kernel<write>");
  return B200LIC_OK;
}

int b200lic_rans_decode(const unsigned* words, const unsigned* chunk_off, unsigned n, unsigned chunk, const int* indexes,
                        const int* cdf, const int* cdf_len, const int* offset, int cdf_stride, int* symbols,
                        b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(words && chunk_off && indexes && cdf && cdf_len && offset && symbols && chunk > 0 && cdf_stride > 2,
               "rans_decode: bad arguments");
  if (n == 0) return B200LIC_OK;
  const unsigned n_chunks = (n + chunk - 1) / chunk;
  rans_decode_kernel<<<(n_chunks + 31) / 32, 32, 0, as_stream(stream)>>>(words, chunk_off, n, chunk, n_chunks, indexes, cdf,
                                                                       cdf_len, offset, cdf_stride, symbols);
  B200_LAUNCH_CHECK("rans_decode_kernel");
  return B200LIC_OK;
}

}  // extern "C"
