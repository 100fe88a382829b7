// MINIMAL STAND-IN for PaddlePaddle's "paddle/extension.h" (Paddle >= 2.1 custom-operator API), containing only the declarations
// paddle_binding/lws_paddle_ops.cc uses.  It exists so that the binding can be syntax- and type-checked in an image without Paddle
// (tests/test_abi.py runs `g++ -fsyntax-only` against it).  With a real Paddle install this directory is NOT on the include path.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace paddle {
enum class DataType { FLOAT32, UINT8, INT32 };
struct Place {};
class Tensor {
 public:
  template <typename T> T* data() const;
  std::vector<int64_t> shape() const;
  Place place() const;
  void* stream() const;  // cudaStream_t of the tensor's device context (Paddle >= 2.3: Tensor::stream())
  bool is_gpu() const;
};
Tensor empty(const std::vector<int64_t>& shape, DataType dtype, const Place& place);
Tensor empty_like(const Tensor& x);
}  // namespace paddle

#define PD_CHECK(cond, ...) \
  do {                      \
    if (!(cond)) throw std::string("PD_CHECK failed: " #cond); \
  } while (0)

struct PdOpBuilder {
  static PdOpBuilder& Make(const char* name);
  PdOpBuilder& Inputs(std::vector<std::string>);
  PdOpBuilder& Outputs(std::vector<std::string>);
  PdOpBuilder& Attrs(std::vector<std::string>);
  template <typename F> PdOpBuilder& SetKernelFn(F);
};
#define PD_BUILD_OP_CAT2(a, b) a##b
#define PD_BUILD_OP_CAT(a, b) PD_BUILD_OP_CAT2(a, b)
#define PD_BUILD_OP(name) static PdOpBuilder& PD_BUILD_OP_CAT(pd_op_builder_##name##_, __LINE__) = PdOpBuilder::Make(#name)
#define PD_BUILD_GRAD_OP(name) static PdOpBuilder& PD_BUILD_OP_CAT(pd_grad_op_builder_##name##_, __LINE__) = PdOpBuilder::Make(#name "_grad")
#define PD_KERNEL(fn) (&fn)
