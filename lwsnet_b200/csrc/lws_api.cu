// Status strings / version / explicit options of the lws_b200 C ABI (include/lws.h).
#include <atomic>
#include <string.h>

#include "lws_common.cuh"

extern "C" const char* lws_status_string(int status) {
  switch (status) {
    case LWS_OK: return "LWS_OK";
    case LWS_ERR_BAD_SHAPE: return "LWS_ERR_BAD_SHAPE";
    case LWS_ERR_BAD_ALIGN: return "LWS_ERR_BAD_ALIGN";
    case LWS_ERR_NULL_PTR: return "LWS_ERR_NULL_PTR";
    case LWS_ERR_WORKSPACE_TOO_SMALL: return "LWS_ERR_WORKSPACE_TOO_SMALL";
    case LWS_ERR_UNSUPPORTED: return "LWS_ERR_UNSUPPORTED";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "LWS_ERR_UNKNOWN";
}

extern "C" const char* lws_version(void) { return "lws_b200 0.2.0 sm_100a"; }

// ---- options: the only process-wide state of the library, changed only through lws_set_option (never read from the environment) ----
namespace lws {
struct OptDesc {
  const char* key;
  int def, lo, hi;
};
static const OptDesc kOpts[OPT_COUNT] = {
    {"conv3d_tc", 1, 0, 1},           // 3D stacks on tcgen05 split-fp16 (1) or on the fp32 FFMA kernels (0)
    {"refine_tc", 1, 0, 1},           // refinement on tcgen05 split-fp16 (1) or on the fp32 FFMA kernels (0)
    {"c8_v1", 0, 0, 1},               // C = 8 stacks: one-plane-per-tile kernel instead of the plane-group kernel
    {"c8_chunk", 0, 0, 1 << 20},      // plane-group kernel: lines per chunk of the work order (0 = whole strips)
    {"k1_dt", 8, 8, 24},              // disparity tile of the direct stage-1 volume kernel (8, 12 or 24)
    {"refine_chain", 0, 0, 4},        // consecutive depthwise-separable blocks per L2-resident chain launch (0 / 1 = one block per launch;
                                      // default off: halves the DRAM traffic of the blocks but measured slower, profiles/r02_chain_ab.txt)
    {"chain_sep_items", 160, 1, 100000},  // chain kernel: queue distance between a producer band and its consumers
    {"warp_div_mode", 0, 0, 1},       // warp coordinate normalisation: 0 = x * fl32(1/c) (Paddle 2.0 scale op), 1 = true division x / c
    {"chain_min_bands", 24, 0, 1 << 20},  // chains are used when the launch has at least this many (pair, band) units
    {"chain_debug", 0, 0, 255},       // TIMING EXPERIMENTS ONLY (results are wrong): 1 = skip dependency waits, 2 = signal without store completion
    {"first_conv", 1, 0, 8},          // first 3D conv (1 -> C): 0 = taps from global memory (r01 kernels); 1 = C = 32 with its tap window
                                      // staged in shared memory (default); 4 / 8 = also C = 8 staged, 4 / 8 rows per thread (slower, A/B)
    {"fused_tail", 0, 0, 1},          // stage tail (softmax regression + upsample + skip + next wflow) as ONE kernel instead of three.
                                      // Off by default: bit-identical but slower at the engine's micro-batch (profiles/r02_tail_ab.txt)
    {"c8_group", 0, 0, 4096},         // C = 8 stacks: depth-first over groups of this many pairs (0 = whole batch layer by layer)
    {"fuse_volume", 1, 0, 2},         // stage 1 through lws_cost_volume_conv3d_stack_f32 (volume built inside the first conv kernel; the
                                      // model reads this, profiles/r02_fuse_volume_ab.txt): 0 = two calls, 1 = fused, 2 = fused with 8-disparity tiles
    {"tz_strips", 5, 0, 7},           // C = 32 stack (profiles/r02_tz_strips_ab.txt).  Bits 0-1, the 32 -> 32 layers: 0 = linear tiling (every tile
                                      // loads its three ky boxes); 1 = tiles walk down y in strips, two of the three boxes stay in the shared-memory
                                      // ring (default); 2 = strips cut into segments dealt round-robin (L2-friendlier, worse balance); 3 = CTA pairs
                                      // (cta_group::2).  Bit 2 (value 4): strips for the closing 32 -> 1 conv too (default)
    {"tz_debug", 0, 0, 15},            // TIMING EXPERIMENTS ONLY (results are wrong): Toeplitz GEMM kernels without A loads (1), epilogue (2), stores (4)
    {"k5_int", 1, 0, 1},              // K5 (rescale + upsample + skip): periodic-tap fast path for the integer scales 2 / 4 / 8 (0 = generic kernel)
    {"fe_tma", 1, 0, 1},              // feature pyramid: stride-1 convs on 16-byte aligned maps read their input tile through one TMA load (0 = per-thread loads)
};
static std::atomic<int> g_opts[OPT_COUNT];
static std::atomic<bool> g_opts_init{false};
static void opts_init() {
  if (g_opts_init.load(std::memory_order_acquire)) return;
  for (int i = 0; i < OPT_COUNT; ++i) g_opts[i].store(kOpts[i].def, std::memory_order_relaxed);
  g_opts_init.store(true, std::memory_order_release);
}
int opt(int id) {
  opts_init();
  return g_opts[id].load(std::memory_order_relaxed);
}
}  // namespace lws

extern "C" int lws_set_option(const char* key, int value) {
  using namespace lws;
  if (!key) return LWS_ERR_NULL_PTR;
  opts_init();
  for (int i = 0; i < OPT_COUNT; ++i)
    if (strcmp(key, kOpts[i].key) == 0) {
      if (value < kOpts[i].lo || value > kOpts[i].hi) return LWS_ERR_BAD_SHAPE;
      g_opts[i].store(value, std::memory_order_relaxed);
      return LWS_OK;
    }
  return LWS_ERR_UNSUPPORTED;
}

extern "C" int lws_get_option(const char* key, int* value) {
  using namespace lws;
  if (!key || !value) return LWS_ERR_NULL_PTR;
  opts_init();
  for (int i = 0; i < OPT_COUNT; ++i)
    if (strcmp(key, kOpts[i].key) == 0) {
      *value = g_opts[i].load(std::memory_order_relaxed);
      return LWS_OK;
    }
  return LWS_ERR_UNSUPPORTED;
}
