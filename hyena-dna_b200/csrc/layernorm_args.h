// Argument blocks of the residual-add + LayerNorm kernels (layernorm.cuh); host-includable.
#pragma once

namespace hy {
namespace ln {

constexpr int kWarps = 8;                 // rows in flight per CTA
constexpr int kMaxV = 8;                  // float4 per lane: D <= 1024 on the register path

struct FwdArgs {
  const float* x;        // (rows, D) mixer / embedding output
  const float* res;      // (rows, D) running residual, or null (first block)
  const float* w;        // (D) LayerNorm weight
  const float* b;        // (D) LayerNorm bias, or null
  float* res_out;        // (rows, D) x + res (the new residual); may be null when res is null (then it equals x)
  float* y;              // (rows, D) normalised output
  float* mean;           // (rows)
  float* rstd;           // (rows)
  long long rows;
  int D;
  float eps;
};

struct BwdArgs {
  const float* dy;       // (rows, D) gradient of the normalised output
  const float* dres;     // (rows, D) gradient arriving on the residual output, or null
  const float* r;        // (rows, D) the residual the forward normalised (res_out)
  const float* w;        // (D)
  const float* mean;     // (rows)
  const float* rstd;     // (rows)
  float* dx;             // (rows, D) gradient of x == gradient of the incoming residual
  float* part;           // (CTAs, 2, D) per-CTA partials of dw, db
  long long rows;
  int D;
};

}  // namespace ln
}  // namespace hy
