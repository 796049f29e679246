// fp8fq_torch.cpp -- thin torch-registered operator layer over the C ABI of libfp8fq.so (include/fp8fq.h).
//
//   TORCH_LIBRARY(fp8fq, ...)  ->  torch.ops.fp8fq.<op>
//
// Each op validates its tensors (TORCH_CHECK -> RuntimeError: CUDA, float32, dense NCHW / channels_last, on the current
// device -- there is no CPU path), allocates the output, takes the stream from at::cuda::getCurrentCUDAStream() and
// makes ONE call into the C ABI; return codes other than 0 become exceptions.  No arithmetic lives here.  This is the
// host binding of the hot path (the ops of the validate forward and of the calibration forward); the ctypes binding in
// _lib.py / ops.py covers the whole ABI, is what the C-ABI tests use, and stays the fallback when this library has not
// been built.  Reference call sites replaced: FPQuantizer.forward (fp8_quantizer.py:194-205), BNFusedHijacker.forward's
// epilogue (quantized_folded_bn.py:39-55), QuantizedBlock.forward's tail (models/resnet_quantized.py:39-46), the
// estimators' min / max (range_estimators.py:61-125) and QuantizationManager.forward's calibration step
// (quantization_manager.py:114-122).
#include <ATen/ATen.h>
#include <c10/cuda/CUDAFunctions.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <vector>

#include "../../include/fp8fq.h"

namespace {

using at::Tensor;
using OptTensor = c10::optional<Tensor>;

constexpr int64_t kMaxFusedElems = int64_t(1) << 31;  // the fused epilogues index with 32 bits (ops.MAX_FUSED_ELEMS)

void* cur_stream() { return static_cast<void*>(c10::cuda::getCurrentCUDAStream().stream()); }

bool is_channels_last(const Tensor& t) {
  if (t.is_contiguous()) return false;
  return (t.dim() == 4 && t.is_contiguous(at::MemoryFormat::ChannelsLast)) ||
         (t.dim() == 5 && t.is_contiguous(at::MemoryFormat::ChannelsLast3d));
}

void require(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (this engine has no CPU path); got device ", t.device());
  TORCH_CHECK(t.scalar_type() == at::kFloat, name, " must be float32; got ", t.scalar_type());
  TORCH_CHECK(t.is_contiguous() || is_channels_last(t), name, " must be contiguous (or dense channels_last)");
  TORCH_CHECK(t.get_device() == c10::cuda::current_device(), name, " lives on ", t.device(),
              " but the current CUDA device is ", (int)c10::cuda::current_device(),
              "; kernels are launched on the current device's current stream (use torch.cuda.device(...))");
}

void require_same_layout(const Tensor& a, const Tensor& b, const char* what) {
  TORCH_CHECK(a.sizes() == b.sizes(), what, ": shape mismatch");
  TORCH_CHECK(a.strides() == b.strides(), what, ": both tensors must use the same memory layout");
}

void require_table(const Tensor& table, int64_t C, double mb, int64_t nb, int64_t sb, const char* what) {
  require(table, what);
  const int64_t stride = fp8fq_table_stride((float)mb, (int)nb, (int)sb);
  TORCH_CHECK(stride > 0, what, ": unsupported format (", stride, ")");
  const int64_t need = stride * (C > 1 ? C : 1);
  TORCH_CHECK(table.numel() >= need, what, ": ", table.numel(), " floats, but ", C, " channel table(s) of this format need ",
              need, " (was it prepared for another format or channel count?)");
}

void require_state(const Tensor& t, int64_t C, const char* what) {
  require(t, what);
  TORCH_CHECK(t.numel() >= C, what, ": ", t.numel(), " entries for ", C, " channel(s)");
}

Tensor out_like(const Tensor& x, const OptTensor& out, const char* what) {
  if (!out.has_value()) return at::empty_like(x);
  require(*out, "out");
  require_same_layout(x, *out, what);
  return *out;
}

const char* err_name(int code) {
  switch (code) {
    case FP8FQ_ERR_BAD_ARG: return "FP8FQ_ERR_BAD_ARG";
    case FP8FQ_ERR_UNSUPPORTED: return "FP8FQ_ERR_UNSUPPORTED";
    case FP8FQ_ERR_ALIGNMENT: return "FP8FQ_ERR_ALIGNMENT";
    case FP8FQ_ERR_WORKSPACE: return "FP8FQ_ERR_WORKSPACE";
    default: return nullptr;
  }
}

void check(int code, const char* what) {
  if (code == 0) return;
  const char* n = err_name(code);
  if (n != nullptr) TORCH_CHECK(false, what, ": ", n);
  TORCH_CHECK(false, what, ": CUDA error ", code);
}

const float* cptr(const Tensor& t) { return t.data_ptr<float>(); }
const float* cptr(const OptTensor& t) { return t.has_value() ? t->data_ptr<float>() : nullptr; }

void rows_hw(const Tensor& x, int64_t Cbn, int64_t* rows, int64_t* hw) {
  TORCH_CHECK(x.dim() >= 2 && x.size(1) == Cbn, "x must be [N, C, ...] with C == number of batch-norm channels");
  int64_t p = 1;
  for (int64_t d = 2; d < x.dim(); ++d) p *= x.size(d);
  *rows = x.size(0) * Cbn;
  *hw = p;
}

// ---- validate path ----------------------------------------------------------------------------------------------
Tensor fake_quant(const Tensor& x, const Tensor& table, int64_t C, double mb, int64_t nb, int64_t sb, const OptTensor& out) {
  require(x, "x");
  require_table(table, C, mb, nb, sb, "table");
  Tensor y = out_like(x, out, "fake_quant");
  const int64_t n = x.numel();
  check(fp8fq_fake_quant_f32(cptr(x), y.data_ptr<float>(), cptr(table), n, C, C > 0 ? n / C : 0, (float)mb, (int)nb, (int)sb,
                             cur_stream()),
        "fp8fq_fake_quant_f32");
  return y;
}

std::vector<Tensor> fake_quant_multi(at::TensorList xs, at::TensorList tables, at::IntArrayRef Cs, double mb, int64_t nb,
                                     int64_t sb) {
  TORCH_CHECK(xs.size() == tables.size() && xs.size() == Cs.size(), "fake_quant_multi: list lengths differ");
  std::vector<Tensor> outs;
  std::vector<fp8fq_tensor_desc> descs(xs.size());
  outs.reserve(xs.size());
  for (size_t i = 0; i < xs.size(); ++i) {
    require(xs[i], "x");
    require_table(tables[i], Cs[i], mb, nb, sb, "table");
    outs.push_back(at::empty_like(xs[i]));
    descs[i] = fp8fq_tensor_desc{cptr(xs[i]), outs[i].data_ptr<float>(), cptr(tables[i]), Cs[i],
                                 Cs[i] > 0 ? xs[i].numel() / Cs[i] : 0};
  }
  check(fp8fq_fake_quant_multi_f32(descs.data(), (int)descs.size(), (float)mb, (int)nb, (int)sb, cur_stream()),
        "fp8fq_fake_quant_multi_f32");
  return outs;
}

OptTensor bn_act_quant(const Tensor& x, const Tensor& bn_scale, const OptTensor& bn_shift, int64_t act, const Tensor& table,
                       double mb, int64_t nb, int64_t sb, int64_t bn_mode, const OptTensor& out) {
  require(x, "x");
  if (x.numel() >= kMaxFusedElems) return c10::nullopt;
  const int64_t Cbn = bn_scale.numel() / (bn_mode == 1 ? 4 : 1);
  int64_t rows, hw;
  rows_hw(x, Cbn, &rows, &hw);
  require_table(table, 1, mb, nb, sb, "table");
  Tensor y = out_like(x, out, "bn_act_quant");
  if (hw == 1 || is_channels_last(x)) {
    check(fp8fq_bn_act_quant_nhwc_f32(cptr(x), y.data_ptr<float>(), cptr(bn_scale), cptr(bn_shift), x.numel() / Cbn, Cbn,
                                      (int)act, (int)bn_mode, cptr(table), (float)mb, (int)nb, (int)sb, cur_stream()),
          "fp8fq_bn_act_quant_nhwc_f32");
  } else {
    check(fp8fq_bn_act_quant_f32(cptr(x), y.data_ptr<float>(), cptr(bn_scale), cptr(bn_shift), rows, hw, Cbn, (int)act,
                                 (int)bn_mode, cptr(table), (float)mb, (int)nb, (int)sb, cur_stream()),
          "fp8fq_bn_act_quant_f32");
  }
  return y;
}

Tensor add_act_quant(const Tensor& a, const Tensor& b, int64_t act, const Tensor& table, double mb, int64_t nb, int64_t sb,
                     const OptTensor& out) {
  require(a, "a");
  require(b, "b");
  require_same_layout(a, b, "add_act_quant");
  require_table(table, 1, mb, nb, sb, "table");
  Tensor y = out_like(a, out, "add_act_quant");
  check(fp8fq_add_act_quant_f32(cptr(a), cptr(b), y.data_ptr<float>(), a.numel(), (int)act, cptr(table), (float)mb, (int)nb,
                                (int)sb, cur_stream()),
        "fp8fq_add_act_quant_f32");
  return y;
}

OptTensor bn_quant_add_act_quant(const Tensor& x, const Tensor& residual, const Tensor& bn_scale, const OptTensor& bn_shift,
                                 int64_t act, const Tensor& table_inner, double mbi, int64_t nbi, int64_t sbi,
                                 const Tensor& table_outer, double mbo, int64_t nbo, int64_t sbo, int64_t bn_mode,
                                 const OptTensor& out) {
  require(x, "x");
  require(residual, "residual");
  require_same_layout(x, residual, "bn_quant_add_act_quant");
  if (x.numel() >= kMaxFusedElems) return c10::nullopt;
  const int64_t Cbn = bn_scale.numel() / (bn_mode == 1 ? 4 : 1);
  int64_t rows, hw;
  rows_hw(x, Cbn, &rows, &hw);
  require_table(table_inner, 1, mbi, nbi, sbi, "table_inner");
  require_table(table_outer, 1, mbo, nbo, sbo, "table_outer");
  Tensor y = out_like(x, out, "bn_quant_add_act_quant");
  if (hw == 1 || is_channels_last(x)) {
    check(fp8fq_bn_quant_add_act_quant_nhwc_f32(cptr(x), cptr(residual), y.data_ptr<float>(), cptr(bn_scale), cptr(bn_shift),
                                                x.numel() / Cbn, Cbn, (int)act, (int)bn_mode, cptr(table_inner), (float)mbi,
                                                (int)nbi, (int)sbi, cptr(table_outer), (float)mbo, (int)nbo, (int)sbo,
                                                cur_stream()),
          "fp8fq_bn_quant_add_act_quant_nhwc_f32");
    return y;
  }
  const int code = fp8fq_bn_quant_add_act_quant_f32(cptr(x), cptr(residual), y.data_ptr<float>(), cptr(bn_scale),
                                                    cptr(bn_shift), rows, hw, Cbn, (int)act, (int)bn_mode, cptr(table_inner),
                                                    (float)mbi, (int)nbi, (int)sbi, cptr(table_outer), (float)mbo, (int)nbo,
                                                    (int)sbo, cur_stream());
  if (code == FP8FQ_ERR_UNSUPPORTED) return c10::nullopt;   // the caller composes the two unfused kernels
  check(code, "fp8fq_bn_quant_add_act_quant_f32");
  return y;
}

// ---- calibration path ---------------------------------------------------------------------------------------------
void minmax(const Tensor& x, bool per_channel, Tensor cur_min, Tensor cur_max, int64_t est_mode, bool initialized,
            double momentum, const Tensor& workspace) {
  require(x, "x");
  const int64_t n = x.numel(), C = per_channel ? x.size(0) : 1;
  require_state(cur_min, C, "cur_min");
  require_state(cur_max, C, "cur_max");
  check(fp8fq_minmax_f32(cptr(x), n, C, C > 0 ? n / C : 0, cur_min.data_ptr<float>(), cur_max.data_ptr<float>(), (int)est_mode,
                         initialized ? 1 : 0, momentum, workspace.data_ptr(), cur_stream()),
        "fp8fq_minmax_f32");
}

void estimate_prepare(const Tensor& x, bool per_channel, Tensor cur_min, Tensor cur_max, int64_t est_mode, bool initialized,
                      double momentum, Tensor maxval_out, double mb, int64_t nb, int64_t sb, Tensor table_out,
                      const Tensor& workspace) {
  require(x, "x");
  const int64_t n = x.numel(), C = per_channel ? x.size(0) : 1;
  require_state(cur_min, C, "cur_min");
  require_state(cur_max, C, "cur_max");
  require_state(maxval_out, C, "maxval_out");
  require_table(table_out, C, mb, nb, sb, "table_out");
  check(fp8fq_estimate_prepare_f32(cptr(x), n, C, C > 0 ? n / C : 0, cur_min.data_ptr<float>(), cur_max.data_ptr<float>(),
                                   (int)est_mode, initialized ? 1 : 0, momentum, maxval_out.data_ptr<float>(), (float)mb,
                                   (int)nb, (int)sb, table_out.data_ptr<float>(), workspace.data_ptr(), cur_stream()),
        "fp8fq_estimate_prepare_f32");
}

bool bn_act_estimate_prepare(const Tensor& x, const Tensor& bn_scale, const OptTensor& bn_shift, int64_t act, int64_t bn_mode,
                             Tensor cur_min, Tensor cur_max, int64_t est_mode, bool initialized, double momentum,
                             const OptTensor& maxval_out, double mb, int64_t nb, int64_t sb, const OptTensor& table_out,
                             const Tensor& workspace) {
  require(x, "x");
  if (x.numel() >= kMaxFusedElems) return false;
  const int64_t Cbn = bn_scale.numel() / (bn_mode == 1 ? 4 : 1);
  int64_t rows, hw;
  rows_hw(x, Cbn, &rows, &hw);
  const bool nhwc = hw == 1 || is_channels_last(x);
  require_state(cur_min, 1, "cur_min");
  require_state(cur_max, 1, "cur_max");
  if (maxval_out.has_value()) require_state(*maxval_out, 1, "maxval_out");
  if (table_out.has_value()) require_table(*table_out, 1, mb, nb, sb, "table_out");
  const int code = fp8fq_bn_act_estimate_prepare_f32(
      cptr(x), nhwc ? x.numel() / Cbn : rows, hw, Cbn, nhwc ? 1 : 0, cptr(bn_scale), cptr(bn_shift), (int)bn_mode, (int)act,
      cur_min.data_ptr<float>(), cur_max.data_ptr<float>(), (int)est_mode, initialized ? 1 : 0, momentum,
      maxval_out.has_value() ? maxval_out->data_ptr<float>() : nullptr, (float)mb, (int)nb, (int)sb,
      table_out.has_value() ? table_out->data_ptr<float>() : nullptr, workspace.data_ptr(), cur_stream());
  if (code == FP8FQ_ERR_UNSUPPORTED) return false;
  check(code, "fp8fq_bn_act_estimate_prepare_f32");
  return true;
}

int64_t abi_version() { return fp8fq_version(); }

}  // namespace

TORCH_LIBRARY(fp8fq, m) {
  m.def("abi_version() -> int", &abi_version);
  m.def("fake_quant(Tensor x, Tensor table, int C, float mantissa_bits, int n_bits, int sign_bits, Tensor? out=None) -> Tensor",
        &fake_quant);
  m.def("fake_quant_multi(Tensor[] xs, Tensor[] tables, int[] Cs, float mantissa_bits, int n_bits, int sign_bits) -> Tensor[]",
        &fake_quant_multi);
  m.def("bn_act_quant(Tensor x, Tensor bn_scale, Tensor? bn_shift, int act, Tensor table, float mantissa_bits, int n_bits, "
        "int sign_bits, int bn_mode=0, Tensor? out=None) -> Tensor?",
        &bn_act_quant);
  m.def("add_act_quant(Tensor a, Tensor b, int act, Tensor table, float mantissa_bits, int n_bits, int sign_bits, "
        "Tensor? out=None) -> Tensor",
        &add_act_quant);
  m.def("bn_quant_add_act_quant(Tensor x, Tensor residual, Tensor bn_scale, Tensor? bn_shift, int act, Tensor table_inner, "
        "float mantissa_bits_inner, int n_bits_inner, int sign_bits_inner, Tensor table_outer, float mantissa_bits_outer, "
        "int n_bits_outer, int sign_bits_outer, int bn_mode=0, Tensor? out=None) -> Tensor?",
        &bn_quant_add_act_quant);
  m.def("minmax(Tensor x, bool per_channel, Tensor(a!) cur_min, Tensor(b!) cur_max, int est_mode, bool initialized, "
        "float momentum, Tensor workspace) -> ()",
        &minmax);
  m.def("estimate_prepare(Tensor x, bool per_channel, Tensor(a!) cur_min, Tensor(b!) cur_max, int est_mode, bool initialized, "
        "float momentum, Tensor(c!) maxval_out, float mantissa_bits, int n_bits, int sign_bits, Tensor(d!) table_out, "
        "Tensor workspace) -> ()",
        &estimate_prepare);
  m.def("bn_act_estimate_prepare(Tensor x, Tensor bn_scale, Tensor? bn_shift, int act, int bn_mode, Tensor(a!) cur_min, "
        "Tensor(b!) cur_max, int est_mode, bool initialized, float momentum, Tensor(c!)? maxval_out, float mantissa_bits, "
        "int n_bits, int sign_bits, Tensor(d!)? table_out, Tensor workspace) -> bool",
        &bn_act_estimate_prepare);
}
