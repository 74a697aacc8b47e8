// Flat fp32 parameter-buffer layouts (state_dict order) of the three networks.
// Reference: core/models/generator.py:17-26, discriminator.py:21-28, forward_model.py:28-60; the key
// order is what nn.Module.state_dict() yields for those nn.Sequential stacks.
#pragma once
#include <stdint.h>

#include "../../include/pigan_b200.h"

namespace pigan {

struct GenLayout {
  int S, H1, H2, P;
  int64_t w1, b1, bn1_w, bn1_b, w2, b2, bn2_w, bn2_b, w3, b3, total;
  // BN buffers (separate contiguous fp32 buffer): rm1, rv1, rm2, rv2
  int64_t rm1, rv1, rm2, rv2, bn_total;
  explicit GenLayout(const PiganDims& d) {
    S = d.spectrum_dim; H1 = d.g_hidden[0]; H2 = d.g_hidden[1]; P = d.param_dim;
    int64_t o = 0;
    w1 = o; o += (int64_t)H1 * S;
    b1 = o; o += H1;
    bn1_w = o; o += H1;
    bn1_b = o; o += H1;
    w2 = o; o += (int64_t)H2 * H1;
    b2 = o; o += H2;
    bn2_w = o; o += H2;
    bn2_b = o; o += H2;
    w3 = o; o += (int64_t)P * H2;
    b3 = o; o += P;
    total = o;
    rm1 = 0; rv1 = H1; rm2 = 2 * (int64_t)H1; rv2 = 2 * (int64_t)H1 + H2; bn_total = 2 * (int64_t)H1 + 2 * H2;
  }
};

struct DiscLayout {
  int S, P, H1, H2, IN;
  int64_t w1, b1, w2, b2, w3, b3, total;
  explicit DiscLayout(const PiganDims& d) {
    S = d.spectrum_dim; P = d.param_dim; H1 = d.d_hidden[0]; H2 = d.d_hidden[1]; IN = S + P;
    int64_t o = 0;
    w1 = o; o += (int64_t)H1 * IN;
    b1 = o; o += H1;
    w2 = o; o += (int64_t)H2 * H1;
    b2 = o; o += H2;
    w3 = o; o += H2;
    b3 = o; o += 1;
    total = o;
  }
};

struct FwdLayout {
  int P, S, Mt, OUT;
  int H[5];
  int64_t w[6], b[6], ln_w[5], ln_b[5], total;
  explicit FwdLayout(const PiganDims& d) {
    P = d.param_dim; S = d.spectrum_dim; Mt = d.metrics_dim; OUT = S + Mt;
    for (int i = 0; i < 5; ++i) H[i] = d.f_hidden[i];
    int64_t o = 0;
    int in = P;
    for (int i = 0; i < 5; ++i) {
      w[i] = o; o += (int64_t)H[i] * in;
      b[i] = o; o += H[i];
      ln_w[i] = o; o += H[i];
      ln_b[i] = o; o += H[i];
      in = H[i];
    }
    w[5] = o; o += (int64_t)OUT * in;
    b[5] = o; o += OUT;
    total = o;
  }
};

inline bool dims_are_default(const PiganDims& d) {
  return d.spectrum_dim == 250 && d.param_dim == 4 && d.metrics_dim == 8 && d.g_hidden[0] == 512 &&
         d.g_hidden[1] == 256 && d.d_hidden[0] == 512 && d.d_hidden[1] == 256 && d.f_hidden[0] == 256 &&
         d.f_hidden[1] == 512 && d.f_hidden[2] == 1024 && d.f_hidden[3] == 512 && d.f_hidden[4] == 256;
}

}  // namespace pigan
