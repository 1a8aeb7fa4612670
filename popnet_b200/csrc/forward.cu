// rtpose_light3d forward: layer plan, weight packing and launch schedule.
//
// Mirrors third_party_methods/lib/network/rtpose_light3d.py of the reference:
//   ResPreprocessNet._forward_impl :201-216, BasicBlock.forward :56-72, make_stages :222-246,
//   rtpose_light3d.__init__ :249-324 (channel tables) and forward :326-356.
// Canonical conv-layer order (what popnet_pack_weights expects, 39 layers for num_stages = 2):
//   0 model0.conv1 | 1-4 model0.layer1.{0,1}.conv{1,2} | 5 layer2.0.conv1 | 6 layer2.0.conv2 |
//   7 layer2.0.downsample.0 | 8 model0.conv2 | 9 + 5*(3*(stage-1) + (branch-1)) + i : model{stage}_{branch}.{3i}
//
// Stage-2 input (reference: torch.cat([paf, heat, depth, feat], 1), 187 channels) is one 192-channel C8P
// buffer laid out [paf 28 + 4 zero | heat 16 | depth 15 + 1 zero | feat 128]; the heads and the second
// average pool write straight into their slices, so the concatenation never happens.
#include <cuda_bf16.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "conv.cuh"

namespace popnet {
namespace {

struct Buf {
  int C, H, W;
  size_t off;             // bf16 elements from the workspace start
  long long plane_stride; // bf16 elements
  int P;                  // (2 + N * (H+1)) * (W+1)
};

// (DA directly behind SA: the fused first layer of the heat-map + depth branches writes 32 consecutive planes)
enum BufId { A112, B112, C112, D56, E56, F56, G56, S2IN, LA, LB, LC, SA, DA, SB_, DB, DC, kNumBufs };

struct Layer {
  int cin, cout, k;          // logical
  int cin_pad, cout_pad;
  int nt, nacc;
  int act;
  int in_buf, in_plane0, out_buf, out_plane0, res_buf;
  int head;                  // 0 none, 1 paf, 2 heat, 3 depth
  int stage;                 // 0 block0, 1, 2
  int remap_s2;              // input channels follow the S2IN layout
  int fuse_layer;            // index of a 1x1 layer whose weights are K-concatenated here (-1 = none); its input is in2_buf
  int in2_buf;
  int mc;                    // launch the cluster-of-two multicast kernel
  int nfuse_layer;           // index of a layer with the SAME input whose output channels are N-concatenated behind this
                             // layer's (-1 = none): one launch computes both, its planes continue into the next buffer
  size_t w_off, shift_off;   // bytes in the packed blob
};

struct Plan {
  Buf bufs[kNumBufs];
  std::vector<Layer> layers;
  size_t blob_bytes = 0, ws_bytes = 0;
  size_t flags_off = 0, flags_bytes = 0;   // progress counters of the layer chain (bytes from the workspace start)
  int K, L, pl, ph, pd;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool make_plan(const PopnetNetConfig& cfg, int batch, Plan& p) {
  if (cfg.operand_dtype != 0 && cfg.operand_dtype != 1) return false;
  if (cfg.input_dim != 1 || cfg.height % 8 || cfg.width % 8 || cfg.height < 16 || cfg.width < 16) return false;
  if (cfg.num_parts < 1 || cfg.num_parts > 15 || cfg.num_limbs < 1 || cfg.num_limbs > 15) return false;
  if (cfg.width / 2 + 2 > kGuard) return false;
  p.K = cfg.num_parts; p.L = cfg.num_limbs;
  p.pl = 32; p.ph = 16; p.pd = 16;
  const int H2 = cfg.height / 2, W2 = cfg.width / 2, H4 = H2 / 2, W4 = W2 / 2, H8 = H4 / 2, W8 = W4 / 2;
  auto setb = [&](int id, int C, int H, int W) { p.bufs[id].C = C; p.bufs[id].H = H; p.bufs[id].W = W; };
  setb(A112, 64, H2, W2); setb(B112, 64, H2, W2); setb(C112, 64, H2, W2);
  setb(D56, 64, H4, W4); setb(E56, 128, H4, W4); setb(F56, 128, H4, W4); setb(G56, 128, H4, W4);
  setb(S2IN, 192, H8, W8);
  setb(LA, 256, H8, W8); setb(LB, 256, H8, W8); setb(LC, 128, H8, W8);
  setb(SA, 128, H8, W8); setb(SB_, 128, H8, W8);
  setb(DA, 128, H8, W8); setb(DB, 64, H8, W8); setb(DC, 64, H8, W8);
  size_t off = 0;
  for (int i = 0; i < kNumBufs; ++i) {
    Buf& b = p.bufs[i];
    b.P = (int)c8p_positions(batch, b.H, b.W);
    b.plane_stride = (long long)(kGuard + align_up((size_t)(batch > 0 ? b.P : 0), kPosRound) + kPosSlack + kGuard) * 8;
    b.off = off;
    off += (size_t)(b.C / 8) * b.plane_stride;
    off = align_up(off, 128);
  }
  p.ws_bytes = off * sizeof(h16);
  p.flags_off = align_up(p.ws_bytes, 256);
  p.flags_bytes = align_up(conv_chain_flag_words(p.bufs[A112].P, 3, 4) * sizeof(unsigned int), 256);
  p.ws_bytes = p.flags_off + p.flags_bytes;

  auto add = [&](int cin, int cout, int k, int nt, int nacc, int act, int in_buf, int in_plane0, int out_buf,
                 int out_plane0, int res_buf, int head, int stage, int remap) {
    Layer l{};
    l.cin = cin; l.cout = cout; l.k = k;
    l.cin_pad = (k == 7) ? cin : (int)align_up(cin, 64);
    l.cout_pad = (int)align_up(cout, nt);
    l.nt = nt; l.nacc = nacc; l.act = act;
    l.in_buf = in_buf; l.in_plane0 = in_plane0; l.out_buf = out_buf; l.out_plane0 = out_plane0; l.res_buf = res_buf;
    l.head = head; l.stage = stage; l.remap_s2 = remap; l.fuse_layer = -1; l.in2_buf = -1; l.nfuse_layer = -1; l.mc = 0;
    p.layers.push_back(l);
  };
  p.layers.clear();
  // block0
  add(1, 64, 7, 64, 0, kActRelu, -1, 0, A112, 0, -1, 0, 0, 0);              // 0 conv1 (stem kernel)
  add(64, 64, 3, 64, 3, kActRelu, A112, 0, B112, 0, -1, 0, 0, 0);           // 1 layer1.0.conv1
  add(64, 64, 3, 64, 3, kActRelu, B112, 0, C112, 0, A112, 0, 0, 0);         // 2 layer1.0.conv2 (+x)
  add(64, 64, 3, 64, 3, kActRelu, C112, 0, B112, 0, -1, 0, 0, 0);           // 3 layer1.1.conv1
  add(64, 64, 3, 64, 3, kActRelu, B112, 0, A112, 0, C112, 0, 0, 0);         // 4 layer1.1.conv2 (+x)
  add(64, 128, 3, 128, 1, kActRelu, D56, 0, E56, 0, -1, 0, 0, 0);           // 5 layer2.0.conv1 (128-position tiles, weights resident)
  add(128, 128, 3, 128, 4, kActRelu, E56, 0, G56, 0, -1, 0, 0, 0);         // 6 layer2.0.conv2, projection shortcut fused:
  p.layers.back().fuse_layer = 7; p.layers.back().in2_buf = D56;            //   relu(bn2(conv2(e)) + bn_d(conv1x1(d))) as ONE K = 9*128 + 64 GEMM
  add(64, 128, 1, 128, 4, kActNone, D56, 0, F56, 0, -1, 0, 0, 0);           // 7 layer2.0.downsample (weights only; never launched)
  add(128, 128, 1, 128, 4, kActRelu, G56, 0, E56, 0, -1, 0, 0, 0);         // 8 conv2
  const int K1 = p.K + 1, L2 = 2 * p.L, L1 = p.L + 1;
  // 28x28 stage layers: 512-position tiles.  At batch 64 the 53,882 positions make 106 tiles (72 % of the 148 SMs), 384-position
  // tiles (POPNET_STAGE_NACC=3) make 141 (95 %); measured identical (1.065 vs 1.067 ms per forward) because the three concurrent
  // branches already fill each other's idle SMs.
  const bool kFuseFirst = p.bufs[DA].off == p.bufs[SA].off + (size_t)(p.bufs[SA].C / 8) * p.bufs[SA].plane_stride &&
                          p.bufs[DA].plane_stride == p.bufs[SA].plane_stride;
  // POPNET_TUNE_MC: the N = 256 stage layers as cluster-of-two multicast kernels with 128-position tiles and double-buffered
  // accumulators.  Validated (tests/test_forward.py) and 20-25 % faster per layer (#11: 56.6 -> 43.5 us in the timeline), but
  // the forward as a whole does not gain (0.980 vs 0.971 ms, same box): the stage is then bounded by the serial heat-map chain
  // and the power cap (the heat-map chain's 128 -> 128 convs as multicast kernels with 256-position tiles: 1.01 ms).
  // Off by default; kept as the basis for cta_group::2 pairs.
  const int kMc = (cfg.tuning & POPNET_TUNE_MC) ? 1 : 0;
  int kStageNacc = 4;
  { const int v = (int)((cfg.tuning >> 2) & 3u); if (v >= 2) kStageNacc = v; }
  for (int s = 1; s <= 2; ++s) {
    const int in0 = (s == 1) ? (p.pl + p.ph + p.pd) / 8 : 0;     // stage 1 reads only the feature planes
    const int cin = (s == 1) ? 128 : 128 + L2 + K1 + L1;
    const int remap = (s == 2);
    // paf branch: 3x3 -> 256, 256, 256, 1x1 -> 128, 1x1 -> 2L
    add(cin, 256, 3, 256, kMc ? 1 : 2, kActLeaky, S2IN, in0, LA, 0, -1, 0, s, remap); p.layers.back().mc = kMc ? 1 : 0;
    add(256, 256, 3, 256, kMc ? 1 : 2, kActLeaky, LA, 0, LB, 0, -1, 0, s, 0);         p.layers.back().mc = kMc ? 1 : 0;
    add(256, 256, 3, 256, kMc ? 1 : 2, kActLeaky, LB, 0, LA, 0, -1, 0, s, 0);         p.layers.back().mc = kMc ? 1 : 0;
    add(256, 128, 1, 128, kStageNacc, kActLeaky, LA, 0, LC, 0, -1, 0, s, 0);
    add(128, L2, 1, 32, kStageNacc, kActHeadPaf, LC, 0, (s == 1) ? S2IN : -1, 0, -1, 1, s, 0);
    // heat-map branch: 3x3 -> 128 x4, 3x3 -> K+1
    add(cin, 128, 3, 128, kStageNacc, kActLeaky, S2IN, in0, SA, 0, -1, 0, s, remap);
    // the first convs of the heat-map and the depth branch read the same input: ONE N = 256 launch (full-rate MMAs, the
    // input tile staged once) whose output planes 0-15 are SA and 16-31 are DA (the buffer right behind it)
    if (kFuseFirst) {
      Layer& f = p.layers.back();
      f.nfuse_layer = (int)p.layers.size() + 4;       // the depth branch's first conv (weights only; never launched)
      f.nt = 256; f.nacc = kMc ? 1 : 2; f.cout_pad = 256; f.mc = kMc ? 1 : 0;
    }
    add(128, 128, 3, 128, kStageNacc, kActLeaky, SA, 0, SB_, 0, -1, 0, s, 0);
    add(128, 128, 3, 128, kStageNacc, kActLeaky, SB_, 0, SA, 0, -1, 0, s, 0);
    add(128, 128, 3, 128, kStageNacc, kActLeaky, SA, 0, SB_, 0, -1, 0, s, 0);
    add(128, K1, 3, 16, kStageNacc, kActHeadHeat, SB_, 0, (s == 1) ? S2IN : -1, p.pl / 8, -1, 2, s, 0);
    // depth branch: 3x3 -> 128, 64, 64, 64, 3x3 -> L+1
    add(cin, 128, 3, 128, kStageNacc, kActLeaky, S2IN, in0, DA, 0, -1, 0, s, remap);
    add(128, 64, 3, 64, kStageNacc, kActLeaky, DA, 0, DB, 0, -1, 0, s, 0);
    add(64, 64, 3, 64, kStageNacc, kActLeaky, DB, 0, DC, 0, -1, 0, s, 0);
    add(64, 64, 3, 64, kStageNacc, kActLeaky, DC, 0, DB, 0, -1, 0, s, 0);
    add(64, L1, 3, 16, kStageNacc, kActHeadPaf, DB, 0, (s == 1) ? S2IN : -1, (p.pl + p.ph) / 8, -1, 3, s, 0);
  }
  size_t boff = 0;
  for (Layer& l : p.layers) {
    size_t wbytes = (l.k == 7) ? (size_t)64 * 64 * sizeof(h16) : (size_t)l.k * l.k * l.cin_pad * l.cout_pad * sizeof(h16);
    if (l.fuse_layer >= 0) wbytes += (size_t)64 * l.cout_pad * sizeof(h16);   // one extra 64-channel 1x1 chunk
    l.w_off = boff; boff = align_up(boff + wbytes, 256);
    l.shift_off = boff; boff = align_up(boff + (size_t)l.cout_pad * sizeof(float), 256);
  }
  p.blob_bytes = boff;
  return true;
}

// S2IN channel -> channel of the reference's torch.cat([paf, heat, depth, feat]) or -1 for padding
int s2_to_ref(const Plan& p, int c) {
  const int L2 = 2 * p.L, K1 = p.K + 1, L1 = p.L + 1;
  if (c < p.pl) return c < L2 ? c : -1;
  c -= p.pl;
  if (c < p.ph) return c < K1 ? L2 + c : -1;
  c -= p.ph;
  if (c < p.pd) return c < L1 ? L2 + K1 + c : -1;
  c -= p.pd;
  return L2 + K1 + L1 + c;
}

h16* buf_ptr(void* ws, const Buf& b, int plane0) {
  return static_cast<h16*>(ws) + b.off + (size_t)plane0 * b.plane_stride + (size_t)kGuard * 8;
}

}  // namespace
}  // namespace popnet

using namespace popnet;

extern "C" int popnet_num_conv_layers(const PopnetNetConfig* cfg) {
  Plan p;
  if (!cfg || !make_plan(*cfg, 1, p)) return POPNET_ERR_UNSUPPORTED;
  return (int)p.layers.size();
}

extern "C" size_t popnet_packed_weight_bytes(const PopnetNetConfig* cfg) {
  Plan p;
  if (!cfg || !make_plan(*cfg, 1, p)) return 0;
  return p.blob_bytes;
}

extern "C" size_t popnet_workspace_bytes(const PopnetNetConfig* cfg, int batch) {
  Plan p;
  if (!cfg || batch < 1 || !make_plan(*cfg, batch, p)) return 0;
  return p.ws_bytes;
}

extern "C" int popnet_pack_weights(const PopnetNetConfig* cfg, const PopnetConvHost* layers, int num_layers,
                                   void* packed_dev, size_t packed_bytes, void* stream) {
  Plan p;
  if (!cfg || !layers || !packed_dev) return POPNET_ERR_INVALID_ARG;
  if (!make_plan(*cfg, 1, p)) return POPNET_ERR_UNSUPPORTED;
  if (num_layers != (int)p.layers.size() || packed_bytes < p.blob_bytes) return POPNET_ERR_INVALID_ARG;
  std::vector<unsigned char> blob(p.blob_bytes, 0);
  for (size_t li = 0; li < p.layers.size(); ++li) {
    const Layer& l = p.layers[li];
    const PopnetConvHost& h = layers[li];
    if (h.cout != l.cout || h.cin != l.cin || h.ksize != l.k || !h.weight_host || !h.scale_host || !h.shift_host)
      return POPNET_ERR_INVALID_ARG;
    float* shift = reinterpret_cast<float*>(blob.data() + l.shift_off);
    for (int n = 0; n < l.cout; ++n) shift[n] = h.shift_host[n];
    const PopnetConvHost* h2 = nullptr;                  // N-concatenated sibling layer (same input)
    if (l.nfuse_layer >= 0) {
      const Layer& f = p.layers[l.nfuse_layer];
      h2 = &layers[l.nfuse_layer];
      if (f.cin != l.cin || f.k != l.k || f.remap_s2 != l.remap_s2 || l.cout + f.cout > l.cout_pad || !h2->weight_host)
        return POPNET_ERR_UNSUPPORTED;
      for (int n = 0; n < f.cout; ++n) shift[l.cout + n] = h2->shift_host[n];
    }
    if (l.k == 7) {                                      // stem: [k8 = kernel row (8th is zero)][cout][8 = column slot]
      h16* w = reinterpret_cast<h16*>(blob.data() + l.w_off);
      for (int n = 0; n < 64; ++n)
        for (int ry = 0; ry < 8; ++ry)
          for (int rx = 0; rx < 8; ++rx) {               // column slot rx holds input column 2ox-4+rx = kernel column rx-1
            const float v = (ry < 7 && rx >= 1) ? h.weight_host[(size_t)n * 49 + ry * 7 + (rx - 1)] * h.scale_host[n] : 0.f;
            w[(ry * 64 + n) * 8 + rx] = f2h16(v, cfg->operand_dtype);
          }
      continue;
    }
    h16* w = reinterpret_cast<h16*>(blob.data() + l.w_off);
    const int taps = l.k * l.k, k8 = l.cin_pad / 8, ntiles = l.cout_pad / l.nt;
    for (int ti = 0; ti < ntiles; ++ti)
      for (int t = 0; t < taps; ++t)
        for (int g = 0; g < k8; ++g)
          for (int nn = 0; nn < l.nt; ++nn)
            for (int j = 0; j < 8; ++j) {
              const int n = ti * l.nt + nn, ci = g * 8 + j;
              const int cref = l.remap_s2 ? s2_to_ref(p, ci) : (ci < l.cin ? ci : -1);
              float v = 0.f;
              if (n < l.cout && cref >= 0) v = h.weight_host[((size_t)n * l.cin + cref) * taps + t] * h.scale_host[n];
              else if (h2 && n >= l.cout && n - l.cout < p.layers[l.nfuse_layer].cout && cref >= 0)
                v = h2->weight_host[((size_t)(n - l.cout) * l.cin + cref) * taps + t] * h2->scale_host[n - l.cout];
              w[((((size_t)ti * taps + t) * k8 + g) * l.nt + nn) * 8 + j] = f2h16(v, cfg->operand_dtype);
            }
    if (l.fuse_layer >= 0) {                             // K-concatenated projection shortcut (cin 64, 1x1, own BN scale)
      const Layer& f = p.layers[l.fuse_layer];
      const PopnetConvHost& hf = layers[l.fuse_layer];
      if (f.cin != 64 || f.cout != l.cout || f.k != 1 || ntiles != 1) return POPNET_ERR_UNSUPPORTED;
      h16* we = w + (size_t)taps * k8 * l.nt * 8;
      for (int g = 0; g < 8; ++g)
        for (int nn = 0; nn < l.nt; ++nn)
          for (int j = 0; j < 8; ++j) {
            const int ci = g * 8 + j;
            we[((size_t)g * l.nt + nn) * 8 + j] =
                f2h16(nn < l.cout ? hf.weight_host[(size_t)nn * f.cin + ci] * hf.scale_host[nn] : 0.f, cfg->operand_dtype);
          }
      for (int n = 0; n < l.cout; ++n) shift[n] += hf.shift_host[n];
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  POPNET_CUDA_TRY(cudaMemcpyAsync(packed_dev, blob.data(), p.blob_bytes, cudaMemcpyHostToDevice, st));
  POPNET_CUDA_TRY(cudaStreamSynchronize(st));            // `blob` is freed on return
  return POPNET_OK;
}

namespace {
// Two auxiliary streams + fork/join events per (device, CALLER stream), created on first use and kept for the life of the
// process (plumbing state; no results live here).  Keyed by the caller's stream so that forwards issued on different
// streams (two batches in flight) do not serialise on each other's branch streams.
struct AuxStreams {
  int device;
  cudaStream_t owner;
  cudaStream_t s[2];
  cudaEvent_t fork, join[2], nfused;
};
std::mutex g_aux_mu;
std::vector<AuxStreams*> g_aux_sets;
AuxStreams* aux_streams(cudaStream_t owner) {
  std::mutex& mu = g_aux_mu;
  std::vector<AuxStreams*>& sets = g_aux_sets;
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  // streams and events belong to a device: the key is (device, caller's stream) -- the default stream handle is the
  // same value on every device
  int on_device = 0;
  AuxStreams* any = nullptr;
  for (AuxStreams* a : sets) {
    if (a->device != device) continue;
    if (a->owner == owner) return a;
    ++on_device;
    any = a;
  }
  if (on_device >= 16) return any;                                 // bounded: share beyond 16 caller streams per device
  AuxStreams* aux = new AuxStreams();
  aux->device = device;
  aux->owner = owner;
  bool ok = true;
  // default priority on purpose: giving the forward's streams the highest priority (so that the overlapped decode of the
  // previous batch only gets idle SMs) was measured 3 % SLOWER per step (1.110 vs 1.082 ms, same box, A/B/A/B)
  for (int i = 0; i < 2; ++i) ok &= cudaStreamCreateWithFlags(&aux->s[i], cudaStreamNonBlocking) == cudaSuccess;
  ok &= cudaEventCreateWithFlags(&aux->fork, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 2; ++i) ok &= cudaEventCreateWithFlags(&aux->join[i], cudaEventDisableTiming) == cudaSuccess;
  ok &= cudaEventCreateWithFlags(&aux->nfused, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { delete aux; return nullptr; }
  sets.push_back(aux);
  return aux;
}
}  // namespace

extern "C" int popnet_release_streams(void) {
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return 0;
  std::lock_guard<std::mutex> lock(g_aux_mu);
  int released = 0;
  for (size_t i = 0; i < g_aux_sets.size();) {
    AuxStreams* a = g_aux_sets[i];
    if (a->device != device) { ++i; continue; }
    for (int k = 0; k < 2; ++k) { cudaStreamSynchronize(a->s[k]); cudaStreamDestroy(a->s[k]); cudaEventDestroy(a->join[k]); }
    cudaEventDestroy(a->fork);
    cudaEventDestroy(a->nfused);
    delete a;
    g_aux_sets.erase(g_aux_sets.begin() + (long)i);
    ++released;
  }
  return released;
}

extern "C" int popnet_forward(const PopnetNetConfig* cfg, const void* packed_dev, const float* x, int batch,
                              float* paf, float* heat, float* depth, float* s1_paf, float* s1_heat, float* s1_depth,
                              void* workspace, size_t workspace_bytes, int impl, void* stream) {
  if (!cfg || !packed_dev || !x || !paf || !heat || !depth || !workspace || batch < 1) return POPNET_ERR_INVALID_ARG;
  if (impl != POPNET_FWD_IMPL_TCGEN05 && impl != POPNET_FWD_IMPL_SIMT) return POPNET_ERR_INVALID_ARG;
  Plan p;
  if (!make_plan(*cfg, batch, p)) return POPNET_ERR_UNSUPPORTED;
  if (workspace_bytes < p.ws_bytes) return POPNET_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned char* blob = static_cast<const unsigned char*>(packed_dev);

  // Zig-zag tile order through the 112 x 112 block (stem forward, layer 1 backward, layer 2 forward, ...): a layer starts on
  // the positions its producer wrote last, the only part of the 105 MB tensor that is still in the 126 MB L2.  Measured in the
  // bench, same box, A/B/A/B: 62.39 k vs 61.77 k frames/s; each 64 -> 64 layer 3 us shorter.  POPNET_TUNE_NO_ZIGZAG disables.
  const int zigzag = (cfg->tuning & POPNET_TUNE_NO_ZIGZAG) ? 0 : 1;
  const bool chain = (cfg->tuning & POPNET_TUNE_CHAIN) != 0 && impl == POPNET_FWD_IMPL_TCGEN05;
  auto make_conv_args = [&](int li, ConvArgs& a) {
    const Layer& l = p.layers[li];
    const Buf& bi = p.bufs[l.in_buf];
    a = ConvArgs{};
    a.in = buf_ptr(workspace, bi, l.in_plane0);
    a.in_plane_stride = bi.plane_stride;
    a.w = reinterpret_cast<const h16*>(blob + l.w_off);
    a.shift = reinterpret_cast<const float*>(blob + l.shift_off);
    if (l.out_buf >= 0) {
      a.out = buf_ptr(workspace, p.bufs[l.out_buf], l.out_plane0);
      a.out_plane_stride = p.bufs[l.out_buf].plane_stride;
    }
    if (l.res_buf >= 0) {
      a.res = buf_ptr(workspace, p.bufs[l.res_buf], 0);
      a.res_plane_stride = p.bufs[l.res_buf].plane_stride;
    }
    if (l.head) {
      float* s1[4] = {nullptr, s1_paf, s1_heat, s1_depth};
      float* s2[4] = {nullptr, paf, heat, depth};
      a.head_out = (l.stage == 1) ? s1[l.head] : s2[l.head];
    }
    a.P = bi.P; a.Hs = bi.H + 1; a.Wp = bi.W + 1;
    a.chunks = l.cin_pad / 64;
    if (l.fuse_layer >= 0) {
      a.in2 = buf_ptr(workspace, p.bufs[l.in2_buf], 0);
      a.in2_plane_stride = p.bufs[l.in2_buf].plane_stride;
      a.chunks2 = 1;
    }
    a.a_stages = 2;                       // double-buffered across chunks AND across tiles (persistent kernel)
    a.act = l.act; a.cout = l.nfuse_layer >= 0 ? l.cout_pad : l.cout; a.cout_pad = l.cout_pad; a.nt = l.nt; a.taps = l.k * l.k;
    a.fmt = cfg->operand_dtype;
    a.mc = l.mc;
    a.reverse = (zigzag && li >= 1 && li <= 4) ? (chain ? 1 : (li & 1)) : 0;
    a.pair = (int)((cfg->tuning >> 4) & 7u);
    a.pair_res = (cfg->tuning & POPNET_TUNE_PAIR_RES) ? 1 : 0;
    a.grid_cap = 148 - 4 * (int)((cfg->tuning >> 9) & 7u);
    a.balance = (cfg->tuning & POPNET_TUNE_BALANCE) ? 1 : 0;
    a.no_prefill = (cfg->tuning & POPNET_TUNE_NO_PREFILL) ? 1 : 0;
    a.cluster2 = ((cfg->tuning & POPNET_TUNE_CLUSTER_ALL) && l.stage >= 1) ? 1 : 0;
  };
  auto run_conv = [&](int li) -> int {
    const Layer& l = p.layers[li];
    ConvArgs a;
    make_conv_args(li, a);
    if (impl == POPNET_FWD_IMPL_SIMT) return launch_conv_simt(a, st);
    // shrink the A staging if the tile does not fit next to two B stages
    int bst = 0;
    if (conv_tc_smem_bytes(a.nt, l.nacc, a.taps, a.a_stages, a.Wp, &bst, false) > 227 * 1024 &&
        conv_tc_smem_bytes(a.nt, l.nacc, a.taps, a.a_stages, a.Wp, &bst, true) > 227 * 1024)
      a.a_stages = 1;
    return launch_conv_tc(a, l.nacc, st);
  };
  auto run_pool = [&](int in_buf, int out_buf, int out_plane0) -> int {
    const Buf& bi = p.bufs[in_buf];
    PoolArgs a{};
    a.in = buf_ptr(workspace, bi, 0); a.in_plane_stride = bi.plane_stride;
    a.out = buf_ptr(workspace, p.bufs[out_buf], out_plane0); a.out_plane_stride = p.bufs[out_buf].plane_stride;
    a.planes = bi.C / 8; a.N = batch; a.H = bi.H; a.W = bi.W; a.fmt = cfg->operand_dtype;
    a.reverse = (zigzag && in_buf == A112 && !chain) ? 1 : 0;      // (after the chain, which ends on the FIRST tiles: forward)
    return launch_pool(a, st);
  };
#define POPNET_TRY(expr) do { int _rc = (expr); if (_rc != POPNET_OK) return _rc; } while (0)
  if (chain)          // (in front of the stem: nothing sits between two kernels, so programmatic dependent launch stays intact)
    POPNET_CUDA_TRY(cudaMemsetAsync(static_cast<unsigned char*>(workspace) + p.flags_off, 0, p.flags_bytes, st));
  {
    const Layer& l = p.layers[0];
    StemArgs a{};
    a.x = x; a.w = reinterpret_cast<const h16*>(blob + l.w_off); a.shift = reinterpret_cast<const float*>(blob + l.shift_off);
    a.out = buf_ptr(workspace, p.bufs[A112], 0); a.out_plane_stride = p.bufs[A112].plane_stride;
    a.N = batch; a.H = cfg->height; a.W = cfg->width; a.fmt = cfg->operand_dtype;
    POPNET_TRY(launch_stem(a, st));
  }
  bool chained = false;
  if (chain) {
    // layers 1-4 (the four 64 -> 64 convolutions at 112 x 112) as ONE launch: a spatial pipeline of four CTA slices whose
    // tensors travel through the L2 (conv_kernels.cu, "CHAIN").  All four walk the tiles from the last to the first: the
    // stem wrote the last tiles last.  Falls back to one launch per layer for geometries the chain does not take
    // (a row pitch above the tile size).
    ConvArgs ca[4];
    for (int li = 1; li <= 4; ++li) make_conv_args(li, ca[li - 1]);
    unsigned int* flags = reinterpret_cast<unsigned int*>(static_cast<unsigned char*>(workspace) + p.flags_off);
    const int rc = launch_conv_chain(ca, 4, p.layers[1].nacc, flags, st);
    if (rc == POPNET_OK) chained = true;
    else if (rc != POPNET_ERR_UNSUPPORTED) return rc;
  }
  if (!chained)
    for (int li = 1; li <= 4; ++li) POPNET_TRY(run_conv(li));
  POPNET_TRY(run_pool(A112, D56, 0));
  POPNET_TRY(run_conv(5));
  POPNET_TRY(run_conv(6));                 // includes the projection shortcut (layer 7) as an extra K chunk
  POPNET_TRY(run_conv(8));
  POPNET_TRY(run_pool(E56, S2IN, (p.pl + p.ph + p.pd) / 8));
  // The three branches of a stage are independent 5-conv chains on the same input: run the heat-map and
  // depth branches on two auxiliary streams so that their CTAs fill the SMs the (two-wave) PAF branch
  // leaves idle.  Fork/join through events keeps the whole forward capturable in a CUDA graph.
  // (Measured alternative, rejected: giving each chain a fixed SM partition -- e.g. 80/38/30 persistent CTAs -- so
  // that all three run side by side from the start: 1.13 ms per forward instead of 0.97; the CTA-time of the stage
  // layers is the same either way and the shared-queue schedule below is already work-conserving.)
  AuxStreams* aux = aux_streams(st);
  if (!aux) return POPNET_ERR_CUDA;
  for (int s = 1; s <= 2; ++s) {
    const int base = 9 + 15 * (s - 1);
    POPNET_CUDA_TRY(cudaEventRecord(aux->fork, st));
    for (int b = 0; b < 2; ++b) POPNET_CUDA_TRY(cudaStreamWaitEvent(aux->s[b], aux->fork, 0));
    cudaStream_t main_st = st;
    const bool fused_first = p.layers[base + 5].nfuse_layer >= 0;
    // (enqueue order PAF, heat-map, depth; heat-map / depth first, or the auxiliary streams at high priority, measured
    //  0.5 - 4 % slower)
    for (int b = 0; b < 3; ++b) {
      st = (b == 0) ? main_st : aux->s[b - 1];
      for (int i = 0; i < 5; ++i) {
        if (b == 2 && i == 0 && fused_first) {           // computed by the heat-map branch's fused first launch
          POPNET_CUDA_TRY(cudaStreamWaitEvent(st, aux->nfused, 0));
          continue;
        }
        POPNET_TRY(run_conv(base + 5 * b + i));
        if (b == 1 && i == 0 && fused_first) POPNET_CUDA_TRY(cudaEventRecord(aux->nfused, st));
      }
    }
    st = main_st;
    for (int b = 0; b < 2; ++b) {
      POPNET_CUDA_TRY(cudaEventRecord(aux->join[b], aux->s[b]));
      POPNET_CUDA_TRY(cudaStreamWaitEvent(st, aux->join[b], 0));
    }
  }
#undef POPNET_TRY
  return POPNET_OK;
}

// ------------------------------------------------------------------------------------------------
// bring-up / unit-test hook (not in the public header): run ONE convolution on caller-provided C8P
// buffers, through either kernel.  tests/test_gpu_conv_unit.py drives every instantiated tile shape.
// ------------------------------------------------------------------------------------------------
struct PopnetDebugConv {
  const void* in; long long in_plane_stride;
  const void* w; const float* shift;
  void* out; long long out_plane_stride;
  const void* res; long long res_plane_stride;
  float* head_out;
  int P, Hs, Wp, chunks, a_stages, act, cout, cout_pad, nt, nacc, taps, impl, fmt, dbg;
  long long* probe;
};

extern "C" __attribute__((visibility("default"))) int popnet_debug_conv(const PopnetDebugConv* d, void* stream) {
  if (!d) return POPNET_ERR_INVALID_ARG;
  ConvArgs a{};
  a.in = static_cast<const h16*>(d->in); a.in_plane_stride = d->in_plane_stride;
  a.w = static_cast<const h16*>(d->w); a.shift = d->shift;
  a.out = static_cast<h16*>(d->out); a.out_plane_stride = d->out_plane_stride;
  a.res = static_cast<const h16*>(d->res); a.res_plane_stride = d->res_plane_stride;
  a.head_out = d->head_out;
  a.P = d->P; a.Hs = d->Hs; a.Wp = d->Wp; a.chunks = d->chunks; a.a_stages = d->a_stages; a.act = d->act;
  a.cout = d->cout; a.cout_pad = d->cout_pad; a.nt = d->nt; a.taps = d->taps; a.fmt = d->fmt; a.dbg = d->dbg; a.probe = d->probe;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d->impl == POPNET_FWD_IMPL_SIMT ? launch_conv_simt(a, st) : launch_conv_tc(a, d->nacc, st);
}

// Timeline tracing hook (not in the public header; tools/forward_timeline.py).  buf: device array of 4*cap uint64 words,
// initialised by the caller to {~0, ~0, 0, 0} per slot; every conv / stem / pool launch after this call takes the next
// slot.  Call with buf = nullptr to stop; returns the number of slots handed out and copies their tags
// (NT*1000 + NACC*100 + TAPS*10 for conv_tc, 1 = stem, 2 = pool) into tags_out.
extern "C" __attribute__((visibility("default"))) int popnet_debug_trace(unsigned long long* buf, int cap, int* tags_out, int max_tags) {
  const int n = popnet::g_trace_next.load() < popnet::g_trace_cap ? popnet::g_trace_next.load() : popnet::g_trace_cap;
  if (tags_out)
    for (int i = 0; i < n && i < max_tags && i < 1024; ++i) tags_out[i] = popnet::g_trace_tags[i];
  popnet::g_trace_buf = buf;
  popnet::g_trace_cap = buf ? cap : 0;
  popnet::g_trace_next.store(0);
  return n;
}
