// hrb_tf_ops.cc -- TensorFlow custom-op face of libhrb200.so.  SOURCE ONLY: there are no TensorFlow headers in
// the build image or on the GPU box, so this file is neither compiled nor tested here.  It shows
// the binding a HandyRec maintainer would build with
//   g++ -shared -fPIC hrb_tf_ops.cc -o hrb_tf_ops.so $(python -c "import tensorflow as tf; print(' '.join(tf.sysconfig.get_compile_flags()+tf.sysconfig.get_link_flags()))") -L. -lhrb200
// and load with tf.load_op_library("hrb_tf_ops.so") behind handyrec.layers.{CustomEmbedding,SequencePoolingLayer,FM}.
// Every op only forwards device pointers to the C ABI in include/hrb200.h on TF's compute stream.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"

#include "../../include/hrb200.h"

namespace tf = tensorflow;

static void* HrbStream(tf::OpKernelContext* ctx) {
  return *reinterpret_cast<void**>(ctx->op_device_context()->stream()->implementation()->GpuStreamMemberHack());
}
#define HRB_TF_CHECK(ctx, call)                                                                  \
  do {                                                                                           \
    int rc_ = (call);                                                                            \
    OP_REQUIRES(ctx, rc_ == HRB_OK, tf::errors::InvalidArgument(hrb_status_str(rc_), ": ", hrb_last_error())); \
  } while (0)

// ---- handyrec/layers/tools.py:87-101 : gather + tiled mask ------------------------------------------------
REGISTER_OP("HrbEmbedding")
    .Input("table: float").Input("ids: int32").Attr("mask_zero: bool = false")
    .Output("out: float").Output("mask: uint8")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      tf::shape_inference::ShapeHandle out;
      TF_RETURN_IF_ERROR(c->Concatenate(c->input(1), c->Vector(c->Dim(c->input(0), 1)), &out));
      c->set_output(0, out);
      c->set_output(1, out);
      return tf::Status::OK();
    });
class HrbEmbeddingOp : public tf::OpKernel {
 public:
  explicit HrbEmbeddingOp(tf::OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("mask_zero", &mask_zero_)); }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& table = ctx->input(0);
    const tf::Tensor& ids = ctx->input(1);
    tf::TensorShape shape = ids.shape();
    shape.AddDim(table.dim_size(1));
    tf::Tensor *out = nullptr, *mask = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, shape, &out));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, shape, &mask));
    HRB_TF_CHECK(ctx, hrb_embedding_fwd(table.flat<float>().data(), table.dim_size(0), (int32_t)table.dim_size(1),
                                        ids.flat<tf::int32>().data(), ids.NumElements(), out->flat<float>().data(),
                                        mask_zero_ ? mask->flat<tf::uint8>().data() : nullptr, nullptr, HrbStream(ctx)));
  }
 private:
  bool mask_zero_;
};
REGISTER_KERNEL_BUILDER(Name("HrbEmbedding").Device(tf::DEVICE_GPU), HrbEmbeddingOp);

// ---- handyrec/layers/sequence.py:26-46 : masked pooling ---------------------------------------------------
REGISTER_OP("HrbSeqPool")
    .Input("x: float").Input("mask: uint8").Attr("method: int").Output("out: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->MakeShape({c->Dim(c->input(0), 0), 1, c->Dim(c->input(0), 2)}));
      return tf::Status::OK();
    });
class HrbSeqPoolOp : public tf::OpKernel {
 public:
  explicit HrbSeqPoolOp(tf::OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("method", &method_)); }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({x.dim_size(0), 1, x.dim_size(2)}), &out));
    HRB_TF_CHECK(ctx, hrb_seq_pool_fwd(x.flat<float>().data(), ctx->input(1).flat<tf::uint8>().data(), x.dim_size(0),
                                       (int32_t)x.dim_size(1), (int32_t)x.dim_size(2), method_, out->flat<float>().data(), HrbStream(ctx)));
  }
 private:
  int method_;
};
REGISTER_KERNEL_BUILDER(Name("HrbSeqPool").Device(tf::DEVICE_GPU), HrbSeqPoolOp);

// ---- handyrec/layers/interaction.py:26-39 : FM -------------------------------------------------------------
REGISTER_OP("HrbFm").Input("x: float").Input("w: float").Input("w0: float").Output("out: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->MakeShape({c->Dim(c->input(0), 0), 1}));
      return tf::Status::OK();
    });
class HrbFmOp : public tf::OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({x.dim_size(0), 1}), &out));
    HRB_TF_CHECK(ctx, hrb_fm_fwd(x.flat<float>().data(), x.dim_size(1) * x.dim_size(2), x.dim_size(0), (int32_t)x.dim_size(1),
                                 (int32_t)x.dim_size(2), ctx->input(1).flat<float>().data(), ctx->input(2).flat<float>().data(),
                                 out->flat<float>().data(), nullptr, HrbStream(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("HrbFm").Device(tf::DEVICE_GPU), HrbFmOp);
// Gradients are registered from Python with @tf.RegisterGradient("HrbFm") etc., forwarding to
// hrb_fm_bwd / hrb_seq_pool_bwd / hrb_embedding_bwd_dense in the same way.
