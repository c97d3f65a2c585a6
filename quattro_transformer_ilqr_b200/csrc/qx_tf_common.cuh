// Packed weight layout of the Quattro transformer (shared by host-side size queries and kernels).
//
// The wire format is the reference's state_dict (transformer_model.py:97-116; key list in
// SURVEY.md 8a-T); the Python host packs it once, in this order, into one float buffer:
//   x_mean[n] x_std[n] u_mean[c] u_std[c]
//   state_embed.{weight[d][n], bias[d]}  control_embed.{weight[d][c], bias[d]}
//   target_embedding[T][d]  pos_encoder.pe[L][d]  (first L = (N+1)+P+T rows)
//   per layer: in_proj_weight[3d][d] in_proj_bias[3d] out_proj.weight[d][d] out_proj.bias[d]
//              linear1.weight[ff][d] linear1.bias[ff] linear2.weight[d][ff] linear2.bias[d]
//              norm1.weight[d] norm1.bias[d] norm2.weight[d] norm2.bias[d]
//   output_linear.{weight[c][d], bias[c]}
// Every tensor starts on a 32-float (128-B) boundary so kernels may use vector / bulk loads.
#pragma once
#include "qx_common.cuh"

namespace qx {

struct TfLayout {
  int n, c, d, ff, T, P, L, layers;
  size_t x_mean, x_std, u_mean, u_std, ws, bs, wc, bc, tgt, pe, layer0, layer_stride, wo, bo, total;
  // offsets inside a layer
  size_t l_inw, l_inb, l_outw, l_outb, l_w1, l_b1, l_w2, l_b2, l_n1w, l_n1b, l_n2w, l_n2b;

  __host__ __device__ explicit TfLayout(const QxTfConfig& g) {
    n = g.state_dim; c = g.control_dim; d = g.d_model; ff = g.dim_ff; T = g.target_len; P = g.prompt_len;
    L = g.horizon + 1 + P + T; layers = g.num_layers;
    size_t o = 0;
    auto take = [&o](size_t count) { const size_t at = o; o = (o + count + 31) / 32 * 32; return at; };
    x_mean = take(n); x_std = take(n); u_mean = take(c); u_std = take(c);
    ws = take((size_t)d * n); bs = take(d);
    wc = take((size_t)d * c); bc = take(d);
    tgt = take((size_t)T * d);
    pe = take((size_t)L * d);
    layer0 = o;
    o = 0;
    l_inw = take((size_t)3 * d * d); l_inb = take(3 * d);
    l_outw = take((size_t)d * d); l_outb = take(d);
    l_w1 = take((size_t)ff * d); l_b1 = take(ff);
    l_w2 = take((size_t)d * ff); l_b2 = take(d);
    l_n1w = take(d); l_n1b = take(d); l_n2w = take(d); l_n2b = take(d);
    layer_stride = o;
    o = layer0 + layer_stride * layers;
    wo = take((size_t)c * d); bo = take(c);
    total = o;
  }
};

}  // namespace qx
