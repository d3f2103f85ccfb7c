// mlp.cuh -- launch interface of the FC network kernels (mlp.cu), shared with search.cu.
#pragma once
#include "common.cuh"

namespace eaz {

struct NetDesc {
  int D, H, A;
  const float* w[4][3];
  const float* b[4][3];
  const uint8_t* bset;
  int hash_bits, hash_io, hash_dim;
  float max_u, novelty_scale;
};
int make_net_desc(const eaz_fc_params* net, const EnvDesc* env, NetDesc* d);

// Where the observation of row b comes from.
struct MlpSource {
  const uint8_t* dense;       // bool [B,D] or null
  const uint8_t* compact;     // compact env states or null
  const int32_t* node_index;  // optional [B]: row b's state is compact[(node_index[b]*B + b)*S]; null: compact[b*S]
  const uint8_t* ds_seen;     // optional DeepSea per-cell "seen" table [D] (replaces hashing the one-hot row)
  const int32_t* cell_index;  // optional DeepSea observation cell of row b's leaf state (written by the tree kernel)
  // tile flags (common.cuh): wait until the tree kernel has bumped tile_done[tile] to tree_epoch * rows_in_tile instead of a
  // grid-wide PDL wait (tree_epoch == 0: PDL wait); bump mlp_done[tile] once per head when the outputs are visible
  const int* tile_done;
  int tree_epoch;
  int* mlp_done;
  int many_trees;  // the caller-visible batch is large (kFlagManyTrees): other sub-batch streams run beside this launch
};

struct MlpOutputs {
  float* logits[2];  // [B,A] for EAZ_HEAD_EXPLOIT / EAZ_HEAD_EXPLORE
  float* value;      // [B]
  float* ube;        // [B]
  float* novelty;    // [B]
};

// Pre-split (3xTF32 hi/lo), pre-tiled weight images for the tensor-core path (mlp_tensor.cu):
// img[head][layer] = K-chunk images in the canonical UMMA layout of umma.cuh.
struct TensorWeights {
  const uint32_t* img[4][3];
  int k1pad;
  const void* h1[4];  // one-hot observations (DeepSea): pre-activated, pre-split layer-1 rows per cell (mlp_gather.cu)
  const void* w2_ck16[4];  // one-hot observations: the W2 images once more in K = 16 chunks (persistent search kernel, psearch.cuh)
  // Range guard of the scaled 3xFP16 split (numeric status block at the end of the image buffer):
  const float* wscale;     // [4][3] power-of-two scale the weight image of (head, layer) was multiplied with, chosen from max |w|
  uint32_t* num_flags;     // sticky bits: EAZ_NUM_* below (read back by eaz_search_numeric_status)
};
constexpr uint32_t kNumWeightsNonFinite = 1u;   // a weight matrix holds inf / nan, or max |w| > 2^20 (no power-of-two scale fits)
constexpr uint32_t kNumActSaturated = 2u;       // a hidden activation exceeded the fp16 range of the scaled split (|h| * 16 > 65504)
constexpr float kActScale = 16.0f;              // activations are multiplied by 16 before the hi / lo split: |h| < 4094 representable
struct NumStatus {
  float wscale[4][3];
  uint32_t wmax_bits[4][3];
  uint32_t flags;
  uint32_t pad[7];
};
int launch_weight_scales_list(const float* const* w, const int* n, int count, NumStatus* ns, cudaStream_t st);                                   // tile_weights.cu
int launch_tile_weights_f16(const float* W, int K, int N, int Kpad, int Npad, const float* scale, void* out, cudaStream_t st, int chunk_k = 32);  // tile_weights.cu
size_t gather_table_bytes(const NetDesc& net);
int prepare_gather_table(const NetDesc& net, int head, void* buf, uint32_t* num_flags, cudaStream_t st);
int launch_mlp_gather(const NetDesc& net, const EnvDesc& env, const MlpSource& src, const TensorWeights& tw, int B, int heads_mask,
                      const MlpOutputs& out, cudaStream_t stream);
size_t tensor_weights_bytes(const NetDesc& net, const EnvDesc& env);
// fill == false only recomputes the pointers into `buf` (the images written by an earlier call are reused)
int prepare_tensor_weights(const NetDesc& net, const EnvDesc& env, int heads_mask, void* buf, TensorWeights* tw, cudaStream_t st, bool fill = true);
int launch_mlp_tensor(const NetDesc& net, const EnvDesc& env, const MlpSource& src, const TensorWeights& tw, int B, int heads_mask,
                      const MlpOutputs& out, cudaStream_t stream);

// heads_mask: bit h set = evaluate head h.  mode: EAZ_MLP_EXACT | EAZ_MLP_TENSOR (needs `tw`).
int launch_mlp(const NetDesc& net, const EnvDesc& env, const MlpSource& src, int B, int heads_mask, const MlpOutputs& out, int mode,
               cudaStream_t stream, const TensorWeights* tw = nullptr);
int mlp_num_launches(int mode);

// DeepSea: seen[cell] for every one-hot observation (one hash per grid cell).
int launch_deepsea_seen_table(const NetDesc& net, const EnvDesc& env, uint8_t* seen, cudaStream_t stream);

}  // namespace eaz
