// Argument block of the persistent AR decode kernel (ar_decode.cu).
#pragma once
#include "common.cuh"

namespace svanon {

struct ArLayerWeights {
  const float* attn_norm;  // [768]
  const float* wqkv;       // [2304][768]
  const float* wo;         // [768][768]
  const float* ffn_norm;   // [768]
  const float* w1;         // [2304][768]
  const float* w3;         // [2304][768]
  const float* w2;         // [768][2304]
};

constexpr int AR_MAX_BATCH = 4;

struct ArStreamDev {
  float* kc;                    // [12][12][max_seq][64]
  float* vc;
  float* fkc;                   // [4][12][8][64]
  float* fvc;
  float* x_audio;               // [768] cached_new_audio_emb (read at start, rewritten at end)
  const long long* content_id;  // device scalar: content id of this frame (used when cond_row is null)
  const float* cond_row;        // [768] explicit second-token row (offline tail: wait4end rows), or null
  const float* noise;           // [8][1000] Exp(1) tape for this step, or null -> counter-based generator
  int* out_codes;               // [8]
  int pos;                      // sequence position of the first of the two new tokens
  unsigned step;
  unsigned long long seed;
  float temperature;            // this stream's sampling arguments (dual_ar_stream.py:1103-1104 defaults 0.7 / 0.7)
  float top_p;
};

struct ArDecodeArgs {
  ArLayerWeights slow[AR_LAYERS];
  ArLayerWeights fast[AR_FAST_LAYERS];
  const float* norm_w;
  const float* output_w;
  const float* fast_norm_w;
  const float* fast_output_w;
  const float* fast_emb;        // [1000][768]
  const float* codebook_emb;    // [8000][768]
  const float* cond_emb;        // [8192][768]  (ARVCWrapper.embedding)
  const float* rope;            // [2048][32][2]
  const float* fast_rope;       // [8][32][2]
  ArStreamDev s[AR_MAX_BATCH];
  float* x;                     // [2B][768]
  float* h;                     // [2B][768]
  float* q;                     // [2B][768]
  float* g;                     // [2B][2304]
  float* part;                  // [B][12][nsplit][2][66]
  float* logits;                // [B][1024]
  unsigned* barrier;            // [2]: arrival counter, counter value saved across launches; zeroed once
  float* dbg_slow_logits;       // [8192] or null
  float* dbg_hidden;            // [768] or null
  float* dbg_fast_logits;       // [8][1000] or null
  float keep_fraction;          // share of the fast-stack weight lines requested with L2 evict_last (rest: evict_first)
  unsigned long long* prof;     // [8] cycle counters per category (ar_decode_staged.cu Prof), or null
  int max_seq;
  int nsplit;
};

int ar_decode_max_batch();
float grid_barrier_probe(unsigned* bar, int iters, float* scratch, int exchange, int grid, cudaStream_t st);
void launch_ar_decode(const ArDecodeArgs& args, int batch, int grid, cudaStream_t st);
// batch-1 variant with TMA-staged weights (ar_decode_staged.cu)
bool ar_decode_staged_supported(int grid);
void launch_ar_decode_staged(const ArDecodeArgs& args, int grid, cudaStream_t st);

}  // namespace svanon
