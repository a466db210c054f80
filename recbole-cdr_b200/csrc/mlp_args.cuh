// mlp_args.cuh -- argument block shared by the fused gather -> MLP -> loss -> backward -> scatter kernels
// (fused_mlp.cu: fp32 FMA row tiles; tc_mlp.cu: tensor-core row tiles).
#pragma once
#include "xdr_common.cuh"

namespace xdr {

constexpr int kMaxLayers = 3;

struct MlpArgs {
  int n_layers;           // 1..3 Linear layers
  int dims[kMaxLayers + 1];
  const float* W[kMaxLayers];   // [dims[l+1], dims[l]]  (nn.Linear.weight layout)
  const float* b[kMaxLayers];   // [dims[l+1]] or NULL
  float* dW[kMaxLayers];        // accumulated (+=) when backward
  float* db[kMaxLayers];
  int hidden_act;         // activation after every layer but the last
  int last_act;           // activation after the last layer (XDR_ACT_NONE for both users)
  int in_mode;            // 0: x = A_u[idx_u]                       (d0 = dim)
                          // 1: x = [max(A_u[u], B_u[u]) | max(A_i[i], B_i[i])]   (d0 = 2*dim)
  int head;               // 0: MSE against T[idx_u] (d_last = dim)   1: sigmoid + BCE with labels (d_last = 1)
  const float *Au, *Bu, *Ai, *Bi, *T;
  int64_t n_u, n_i;
  int dim;
  const int64_t* idx_u;
  const int64_t* idx_i;
  const float* label;
  int64_t batch;
  int tile_rows;          // rows per tile (<= kMaxTileRows)
  int backward;           // 0: forward only (loss [+ prob]); 1: forward + backward + scatter
  const float* grad_loss; // device scalar (NULL => 1)
  float scale;
  float *dAu, *dBu, *dAi, *dBi, *dT;  // scatter-add destinations (backward)
  float* prob;            // optional [batch] sigmoid output (head 1)
  float* out8;
  int32_t* oob;
};

}  // namespace xdr
