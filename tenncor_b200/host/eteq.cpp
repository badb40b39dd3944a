// eteq.cpp — device-resident holders, the typed_exec switch, graph nodes and factories.
#include "eteq.hpp"

#include <cstdlib>
#include <cstring>

namespace eigen {

void* DeviceRuntimeMemory::allocate(size_t size) {
  cuda::ensure_device();
  void* p = nullptr;
  cuda::check(tcr_alloc(&p, size), "tcr_alloc");
  return p;
}

void DeviceRuntimeMemory::deallocate(void* ptr, size_t) {
  // never throws: holders and plans are also destroyed during interpreter shutdown, after the
  // library's arena bookkeeping may already be gone
  tcr_free(ptr);
}

static RTMemptrT& runtime_slot() {
  static RTMemptrT slot = std::make_shared<DeviceRuntimeMemory>();
  return slot;
}
void set_runtime(RTMemptrT mem) { runtime_slot() = mem ? mem : std::make_shared<DeviceRuntimeMemory>(); }
RTMemptrT get_runtime() { return runtime_slot(); }

void Expirable::expire() {
  if (ptr_ != nullptr) {
    allocator_->deallocate(ptr_, size_);
    ptr_ = nullptr;
    size_ = 0;
    allocator_ = nullptr;
  }
  ttl_ = 0;
}

void Expirable::borrow(RTMemptrT& memory, size_t bytes, size_t ttl) {
  if (nullptr == memory) global::fatal("cannot borrow from null memory");
  if (false == is_expired()) global::throw_err("cannot borrow memory when Expirable is not expired");
  size_ = bytes;
  ptr_ = memory->allocate(size_);
  allocator_ = memory;
  extend_life(ttl);
}

void Expirable::extend_life(size_t ttl) {
  if (nullptr == ptr_) global::fatal("cannot extend ttl of expired Expirable");
  if (ttl_ < ttl) ttl_ = ttl;
}

Observable::Observable(const teq::TensptrsT& args) {
  for (auto& arg : args)
    if (auto f = dynamic_cast<Observable*>(arg.get())) f->subscribe(this);
}

Observable::Observable(const teq::TensptrsT& args, marsh::Maps&& attrs) : Observable(args) { attrs_ = std::move(attrs); }

}  // namespace eigen

namespace cuda {

using namespace teq;

void check(int rc, const char* what) {
  if (rc != TCR_OK) global::fatalf("%s failed (%d): %s", what, rc, tcr_last_error());
}

void ensure_device() {
  static bool ready = false;
  if (ready) return;
  int dev = 0;
  if (const char* lr = std::getenv("LOCAL_RANK")) dev = std::atoi(lr);
  if (const char* d = std::getenv("TCR_DEVICE")) dev = std::atoi(d);
  check(tcr_init(dev), "tcr_init");
  ready = true;
}

void sync() { check(tcr_sync(), "tcr_sync"); }

static int g_gemm_precision = TCR_GEMM_3XTF32;
int gemm_precision() { return g_gemm_precision; }
void set_gemm_precision(int p) {
  if (p != TCR_GEMM_EXACT && p != TCR_GEMM_TF32 && p != TCR_GEMM_3XTF32) global::fatalf("unknown gemm precision %d", p);
  g_gemm_precision = p;
}

void* HostMirror::sync_from(const void* dev, size_t bytes) {
  if (!valid_ || host_.size() != bytes) {
    host_.resize(bytes);
    check(tcr_d2h(host_.data(), dev, bytes), "tcr_d2h");
    sync();
    valid_ = true;
  }
  return host_.data();
}

// ---------------------------------------------------------------- DevSrc
DevSrc::DevSrc(const void* host_data, egen::_GENERATED_DTYPE dtype, Shape shape, bool keep_host)
    : bytes_(shape.n_elems() * egen::type_size(dtype)), keep_host_(keep_host) {
  // staged on the host; the HBM buffer is created at first device use so that graphs can be
  // built (and their shapes / gradients inspected) on a machine without a GPU
  mirror_.host_.assign((const char*)host_data, (const char*)host_data + bytes_);
  mirror_.valid_ = true;
}

DevSrc::~DevSrc() {
  if (dev_) tcr_free(dev_);
  if (staging_) tcr_free(staging_);
}

void DevSrc::prefetch_host(const void* host_data) {
  if (staging_ == nullptr) {
    device_data();
    check(tcr_alloc(&staging_, bytes_), "tcr_alloc");
    sync();  // the block may be recycled from work still queued on the library stream; the copy stream is not ordered with it
  }
  check(tcr_h2d_prefetch(staging_, host_data, bytes_), "tcr_h2d_prefetch");
  staged_ = true;
}

void DevSrc::commit_prefetch() {
  if (!staged_) global::fatal("commit without a prefetched batch");
  check(tcr_prefetch_commit(dev_, staging_, bytes_), "tcr_prefetch_commit");
  staged_ = false;
  mirror_.invalidate();
}

void* DevSrc::data() {
  if (!mirror_.valid_) mirror_.sync_from(dev_, bytes_);
  return mirror_.host_.data();
}

void* DevSrc::device_data() {
  if (dev_ == nullptr) {
    ensure_device();
    check(tcr_alloc(&dev_, bytes_), "tcr_alloc");
    check(tcr_h2d(dev_, mirror_.host_.data(), bytes_), "tcr_h2d");
    sync();  // the pageable staging copy may be released below
    if (!keep_host_) {
      std::vector<char>().swap(mirror_.host_);
      mirror_.valid_ = false;
    }
  }
  return dev_;
}

void DevSrc::assign_host(const void* host_data) {
  if (dev_ == nullptr) {
    mirror_.host_.assign((const char*)host_data, (const char*)host_data + bytes_);
    mirror_.valid_ = true;
    return;
  }
  check(tcr_h2d(dev_, host_data, bytes_), "tcr_h2d");
  if (keep_host_) {
    mirror_.host_.assign((const char*)host_data, (const char*)host_data + bytes_);
    mirror_.valid_ = true;
  } else {
    mirror_.invalidate();
  }
}

void DevSrc::assign_device(const void* dev_data) {
  if (dev_ == nullptr) {
    ensure_device();
    check(tcr_alloc(&dev_, bytes_), "tcr_alloc");
  }
  check(tcr_d2d(dev_, dev_data, bytes_), "tcr_d2d");
  mirror_.invalidate();
}

}  // namespace cuda

namespace cuda {

using namespace teq;

// ---------------------------------------------------------------- DevOp / DevAssign
void DevOp::assign(size_t ttl, eigen::RTMemptrT& runtime) {
  if (data_.is_expired()) data_.borrow(runtime, bytes_, ttl);
  else data_.extend_life(ttl);
  std::vector<const void*> in;
  std::vector<Once<const void*>> onces;
  in.reserve(args_.size());
  onces.reserve(args_.size());
  for (auto& arg : args_) {
    Once<const void*> argdata = arg->device().odata();
    if (nullptr == argdata.get()) global::fatalf("argument %s has no data", arg->to_string().c_str());
    in.push_back(argdata.get());
    onces.push_back(std::move(argdata));
  }
  launch_(data_.get(), in);
  mirror_.invalidate();
}

void* DevOp::ensure_buffer(size_t ttl, eigen::RTMemptrT& runtime) {
  if (data_.is_expired()) data_.borrow(runtime, bytes_, ttl);
  else data_.extend_life(ttl);
  mirror_.invalidate();
  return data_.get();
}

void DevAssign::assign(size_t ttl, eigen::RTMemptrT&) {
  extend_life(ttl);
  auto next_version = arg_->get_meta().state_version() + 1;
  static_cast<eigen::iMutableLeaf*>(ref_)->upversion(next_version);
  void* dst = ref_->device().device_data();
  Once<const void*> src = arg_->device().odata();
  if (nullptr == src.get()) global::fatalf("assign source %s has no data", arg_->to_string().c_str());
  auto dtype = (egen::_GENERATED_DTYPE)ref_->get_meta().type_code();
  check(tcr_assign(op_, dst, src.get(), (int64_t)ref_->shape().n_elems(), dtype), "tcr_assign");
  static_cast<eigen::iEigen&>(ref_->device()).mark_device_dirty();
}

void Device::calc(iTensor& tens, size_t cache_ttl) {
  auto& obs = static_cast<eigen::Observable&>(tens);
  auto& obsdev = static_cast<eigen::iEigen&>(tens.device());
  size_t valid_ttl = obs.nsubs() + cache_ttl;
  // always assign when device is stateless (even if version is the same)
  if (obs.prop_version(max_version_) || nullptr == obsdev.device_data())
    obsdev.assign(std::max<size_t>(1, valid_ttl), memory_);
  else if (false == obsdev.valid_for(valid_ttl))
    obsdev.extend_life(valid_ttl);
}

// ---------------------------------------------------------------- contraction -> GEMM
namespace {
struct RankRun {  // a run of consecutive non-1 ranks that can be walked with one stride
  int64_t extent = 1, stride = 0;
  bool ok = true, empty = true;
  int last = -1;
};

// ranks (ascending) -> single strided index when they are adjacent among the non-1 ranks
RankRun make_run(const Shape& shape, const std::vector<int>& ranks) {
  RankRun run;
  int64_t strides[rank_cap], s = 1;
  for (int r = 0; r < rank_cap; ++r) { strides[r] = s; s *= shape.at(r); }
  int64_t expect = -1;
  for (int r : ranks) {
    if (shape.at(r) == 1) continue;
    if (run.empty) { run.stride = strides[r]; run.empty = false; }
    else if (strides[r] != expect) run.ok = false;
    run.extent *= shape.at(r);
    expect = strides[r] * shape.at(r);
  }
  return run;
}
}  // namespace

bool contract_as_gemm(const Shape& ashape, const Shape& bshape, const eigen::PairVecT<RankT>& pairs, tcr_gemm_desc& d) {
  bool acom[rank_cap] = {false}, bcom[rank_cap] = {false};
  eigen::PairVecT<RankT> sorted(pairs.begin(), pairs.end());
  std::sort(sorted.begin(), sorted.end());
  std::vector<int> ak, bk, am, bn;
  for (auto& p : sorted) {
    acom[p.first] = bcom[p.second] = true;
    ak.push_back(p.first);
    bk.push_back(p.second);
  }
  // the K index must walk both operands in the same order: b's common ranks ascending too
  for (size_t i = 1; i < bk.size(); ++i)
    if (bk[i] < bk[i - 1] && ashape.at(ak[i]) != 1 && ashape.at(ak[i - 1]) != 1) return false;
  for (int r = 0; r < rank_cap; ++r) {
    if (!acom[r]) am.push_back(r);
    if (!bcom[r]) bn.push_back(r);
  }
  RankRun rak = make_run(ashape, ak), rbk = make_run(bshape, bk), ram = make_run(ashape, am), rbn = make_run(bshape, bn);
  if (!rak.ok || !rbk.ok || !ram.ok || !rbn.ok || rak.extent != rbk.extent) return false;
  std::memset(&d, 0, sizeof(d));
  d.m = ram.extent; d.n = rbn.extent; d.k = rak.extent; d.batch = 1;
  d.a_sm = ram.stride; d.a_sk = rak.stride;
  d.b_sk = rbk.stride; d.b_sn = rbn.stride;
  d.c_sn = 1; d.c_sm = d.n;  // out ranks: b-free (fast) then a-free
  return true;
}

void matmul_as_gemm(const Shape& ashape, const Shape& bshape, tcr_gemm_desc& d) {
  std::memset(&d, 0, sizeof(d));
  int64_t K = ashape.at(0), M = ashape.at(1), N = bshape.at(0);
  int64_t batch = 1;
  for (int r = 2; r < rank_cap; ++r) batch *= ashape.at(r);
  d.m = M; d.n = N; d.k = K; d.batch = batch;
  d.a_sm = K; d.a_sk = 1; d.a_sb = M * K;
  d.b_sk = N; d.b_sn = 1; d.b_sb = K * N;
  d.c_sm = N; d.c_sn = 1; d.c_sb = M * N;
}

// ---------------------------------------------------------------- typed_exec
static void shape8(int64_t out[8], const Shape& s) {
  for (int r = 0; r < rank_cap; ++r) out[r] = s.at(r);
}

void typed_exec(egen::_GENERATED_OPCODE opcode, egen::_GENERATED_DTYPE dtype, eigen::EigenptrT& out, Shape outshape,
                const TensptrsT& in, const marsh::iAttributed& attrib) {
  using namespace egen;
  const int64_t n_out = (int64_t)outshape.n_elems();
  const int es = egen::type_size(dtype);
  const size_t out_bytes = (size_t)n_out * es;
  CTensT args;
  for (auto& t : in) args.push_back(t.get());
  auto op = [&](LaunchF f) { out = std::make_shared<DevOp>(out_bytes, args, std::move(f)); };

  switch (opcode) {
    case IDENTITY:
      if (auto red = dynamic_cast<const marsh::Float*>(attrib.get_attr("dp_allreduce"))) {
        // data-parallel gradient exchange (dp.hpp): SUM over ranks, then scale
        double scale = red->val_;
        args.resize(1);
        op([=](void* o, const std::vector<const void*>& a) {
          if (o != a[0]) check(tcr_d2d(o, a[0], out_bytes), "tcr_d2d");  // the planner may run the exchange in place
          check(tcr_allreduce_sum(o, n_out, dtype, scale), "tcr_allreduce_sum");
        });
        break;
      }
      [[fallthrough]];
    case RESHAPE:
      out = std::make_shared<DevRef>(*in[0]);  // eigen::ref (src/operator.cpp:12-15): alias, no data movement
      break;
    case ABS: case NEG: case SIN: case COS: case TAN: case EXP: case LOG: case SQRT: case ROUND: case SIGMOID:
    case TANH: case SQUARE: case CUBE:
      args.resize(1);
      op([=](void* o, const std::vector<const void*>& a) { check(tcr_unary(opcode, a[0], o, n_out, dtype), "tcr_unary"); });
      break;
    case RAND_UNIF:
      op([=](void* o, const std::vector<const void*>& a) {
        eteq::rng_flush();  // (seed, offset) live on the device: graph replays draw fresh numbers
        eteq::rng_advance((uint64_t)n_out);
        check(tcr_rand_unif_stream(a[0], a[1], o, n_out, dtype), "tcr_rand_unif_stream");
      });
      break;
    case REVERSE: {
      uint32_t mask = 0;
      for (RankT r : eigen::unpack_rankset(attrib)) mask |= 1u << r;
      Shape ishape = in[0]->shape();
      op([=](void* o, const std::vector<const void*>& a) {
        int64_t s[8];
        shape8(s, ishape);
        check(tcr_reverse(a[0], o, s, mask, es), "tcr_reverse");
      });
    } break;
    case REDUCE_SUM: case REDUCE_PROD: case REDUCE_MIN: case REDUCE_MAX: {
      uint32_t mask = 0;
      for (RankT r : eigen::unpack_rankset(attrib))
        if (r < rank_cap) mask |= 1u << r;
      Shape ishape = in[0]->shape();
      op([=](void* o, const std::vector<const void*>& a) {
        int64_t s[8];
        shape8(s, ishape);
        check(tcr_reduce(opcode, a[0], o, s, mask, dtype), "tcr_reduce");
      });
    } break;
    case ARGMAX: {
      int return_dim = eigen::unpack_rank(attrib);
      Shape ishape = in[0]->shape();
      op([=](void* o, const std::vector<const void*>& a) {
        int64_t s[8];
        shape8(s, ishape);
        check(tcr_argmax(a[0], o, s, return_dim, dtype), "tcr_argmax");
      });
    } break;
    case PERMUTE: {
      RanksT order = eigen::unpack_ranks(attrib);
      bool visited[rank_cap] = {false};
      if (order.size() > rank_cap) order.resize(rank_cap);
      for (auto r : order) visited[r] = true;
      for (RankT i = 0; i < rank_cap; ++i)
        if (!visited[i]) order.push_back(i);
      Shape ishape = in[0]->shape();
      op([=](void* o, const std::vector<const void*>& a) {
        int64_t s[8];
        int32_t ord[8];
        shape8(s, ishape);
        for (int r = 0; r < 8; ++r) ord[r] = order[r];
        check(tcr_permute(a[0], o, s, ord, es), "tcr_permute");
      });
    } break;
    case EXTEND: {
      Shape ishape = in[0]->shape();
      DimsT bcast = eigen::unpack_extend(ishape, attrib).second;
      op([=](void* o, const std::vector<const void*>& a) {
        int64_t s[8], bc[8];
        shape8(s, ishape);
        for (int r = 0; r < 8; ++r) bc[r] = r < (int)bcast.size() ? bcast[r] : 1;
        check(tcr_extend(a[0], o, s, bc, es), "tcr_extend");
      });
    } break;
    case SLICE: {
      auto encoding = eigen::unpack_dimpairs(attrib);
      Shape shape = in[0]->shape();
      int64_t offsets[8], extents[8];
      for (int r = 0; r < 8; ++r) { offsets[r] = 0; extents[r] = shape.at(r); }
      for (size_t i = 0, n = std::min(encoding.size(), (size_t)rank_cap); i < n; ++i) {
        DimT offset = std::min(encoding[i].first, (DimT)(shape.at(i) - 1));
        offsets[i] = offset;
        extents[i] = std::min(encoding[i].second, (DimT)(shape.at(i) - offset));
      }
      auto slist = narrow_shape(shape);
      if (slist.size() > 0 && outshape.compatible_before(shape, slist.size() - 1)) {
        // only the last non-1 rank is cut: zero-copy view at a pointer offset (operator.hpp:231-248)
        RankT lastdim = slist.size() - 1;
        size_t batchsize = shape.n_elems() / shape.at(lastdim);
        out = std::make_shared<DevRef>(*in[0], (size_t)offsets[lastdim] * batchsize * es);
        break;
      }
      std::vector<int64_t> offs(offsets, offsets + 8), exts(extents, extents + 8);
      op([=](void* o, const std::vector<const void*>& a) {
        int64_t s[8];
        shape8(s, shape);
        check(tcr_slice(a[0], o, s, offs.data(), exts.data(), es), "tcr_slice");
      });
    } break;
    case PAD: {
      auto encoding = eigen::unpack_dimpairs(attrib);
      Shape ishape = in[0]->shape();
      std::vector<int64_t> lo(8, 0), hi(8, 0);
      for (size_t i = 0, n = std::min(encoding.size(), (size_t)rank_cap); i < n; ++i) { lo[i] = encoding[i].first; hi[i] = encoding[i].second; }
      op([=](void* o, const std::vector<const void*>& a) {
        int64_t s[8];
        shape8(s, ishape);
        check(tcr_pad(a[0], o, s, lo.data(), hi.data(), es), "tcr_pad");
      });
    } break;
    case STRIDE: {
      DimsT c = eigen::unpack_dims(attrib);
      Shape ishape = in[0]->shape();
      tcr_map_desc d;
      for (int r = 0; r < 8; ++r) {
        d.in_shape[r] = ishape.at(r); d.out_shape[r] = outshape.at(r); d.perm[r] = r;
        d.mul[r] = r < (int)c.size() ? c[r] : 1; d.add[r] = 0; d.div[r] = 1;
        if ((d.out_shape[r] - 1) * d.mul[r] >= d.in_shape[r])
          global::fatalf("stride: output shape %s reads past input %s", outshape.to_string().c_str(), ishape.to_string().c_str());
      }
      op([=](void* o, const std::vector<const void*>& a) { check(tcr_map_copy(a[0], o, &d, es), "tcr_map_copy"); });
    } break;
    case SCATTER: {
      DimsT c = eigen::unpack_dims(attrib);
      Shape ishape = in[0]->shape();
      op([=](void* o, const std::vector<const void*>& a) {
        int64_t s[8], os[8], inc[8];
        shape8(s, ishape);
        shape8(os, outshape);
        for (int r = 0; r < 8; ++r) inc[r] = r < (int)c.size() ? c[r] : 1;
        check(tcr_scatter(a[0], o, s, os, inc, es), "tcr_scatter");
      });
    } break;
    case POW: case SUB: case DIV: case MIN: case MAX: case EQ: case NEQ: case LT: case GT:
      op([=](void* o, const std::vector<const void*>& a) { check(tcr_binary(opcode, a[0], a[1], o, n_out, dtype), "tcr_binary"); });
      break;
    case ADD: case MUL:
      op([=](void* o, const std::vector<const void*>& a) {
        if (a.size() == 2) check(tcr_binary(opcode, a[0], a[1], o, n_out, dtype), "tcr_binary");
        else check(tcr_nnary(opcode, a.data(), (int)a.size(), o, n_out, dtype), "tcr_nnary");
      });
      break;
    case MATMUL: {
      tcr_gemm_desc d;
      std::memset(&d, 0, sizeof(d));
      matmul_as_gemm(in[0]->shape(), in[1]->shape(), d);
      d.dtype = dtype;
      op([=](void* o, const std::vector<const void*>& a) {
        tcr_gemm_desc dd = d;
        dd.precision = dtype == FLOAT ? gemm_precision() : TCR_GEMM_EXACT;
        check(tcr_gemm(a[0], a[1], o, &dd), "tcr_gemm");
      });
      static_cast<DevOp*>(out.get())->set_gemm(d);
    } break;
    case CONTRACT: {
      auto pairs = eigen::unpack_rankpairs(attrib);
      Shape ashape = in[0]->shape(), bshape = in[1]->shape();
      tcr_gemm_desc d;
      std::memset(&d, 0, sizeof(d));
      if (contract_as_gemm(ashape, bshape, pairs, d)) {
        d.dtype = dtype;
        op([=](void* o, const std::vector<const void*>& a) {
          tcr_gemm_desc dd = d;
          dd.precision = dtype == FLOAT ? gemm_precision() : TCR_GEMM_EXACT;
          check(tcr_gemm(a[0], a[1], o, &dd), "tcr_gemm");
        });
        static_cast<DevOp*>(out.get())->set_gemm(d);
      } else {
        std::vector<int32_t> flat;
        for (auto& p : pairs) { flat.push_back(p.first); flat.push_back(p.second); }
        op([=](void* o, const std::vector<const void*>& a) {
          int64_t sa[8], sb[8];
          shape8(sa, ashape);
          shape8(sb, bshape);
          check(tcr_contract(a[0], a[1], o, sa, sb, flat.data(), (int)flat.size() / 2, dtype), "tcr_contract");
        });
      }
    } break;
    case CONV: {
      Shape ishape = in[0]->shape(), kshape = in[1]->shape();
      RanksT order = eigen::unpack_ranks(attrib);
      bool visited[rank_cap] = {false};
      size_t n = std::min(order.size(), (size_t)rank_cap);
      for (size_t i = 0; i < n; ++i) {
        if (visited[order[i]])
          global::fatalf("convolution does not support repeated kernel dimensions: %s", fmts::to_string(order.begin(), order.end()).c_str());
        visited[order[i]] = true;
      }
      for (size_t i = n; i < rank_cap; ++i)
        if (kshape.at(i) > 1)
          global::fatalf("given kernel shape %s, unspecified non-singular kernel dimension %d is undefined", kshape.to_string().c_str(), (int)i);
      order.resize(n);
      for (RankT i = 0; i < rank_cap; ++i)
        if (!visited[i]) order.push_back(i);
      op([=](void* o, const std::vector<const void*>& a) {
        int64_t si[8], sk[8];
        int32_t ord[8];
        shape8(si, ishape);
        shape8(sk, kshape);
        for (int r = 0; r < 8; ++r) ord[r] = order[r];
        check(tcr_conv(a[0], a[1], o, si, sk, ord, dtype), "tcr_conv");
      });
    } break;
    case SELECT:
      op([=](void* o, const std::vector<const void*>& a) { check(tcr_select(a[0], a[1], a[2], o, n_out, dtype), "tcr_select"); });
      break;
    case CONCAT: {
      int axis = eigen::unpack_rank(attrib);
      std::vector<int64_t> shapes;
      for (auto& t : in)
        for (int r = 0; r < 8; ++r) shapes.push_back(t->shape().at(r));
      op([=](void* o, const std::vector<const void*>& a) {
        check(tcr_concat(a.data(), shapes.data(), (int)a.size(), o, axis, es), "tcr_concat");
      });
    } break;
    case ASSIGN: case ASSIGN_ADD: case ASSIGN_SUB: case ASSIGN_MUL: case ASSIGN_DIV:
      if (nullptr == dynamic_cast<eigen::iMutableLeaf*>(in[0].get()))
        global::fatalf("cannot %s to non-variable %s", name_op(opcode).c_str(), in[0]->to_string().c_str());
      out = std::make_shared<DevAssign>(opcode, *in[0], *in[1]);
      break;
    case CAST: {
      auto intype = (egen::_GENERATED_DTYPE)in[0]->get_meta().type_code();
      if (intype == dtype) {
        out = std::make_shared<DevRef>(*in[0]);
        break;
      }
      op([=](void* o, const std::vector<const void*>& a) { check(tcr_cast(a[0], intype, o, dtype, n_out), "tcr_cast"); });
    } break;
    default:
      global::fatal("unknown opcode");
  }
}

}  // namespace cuda

// ======================================================================== eteq
namespace eteq {

using namespace teq;

static size_t g_lastvers = 0;
size_t get_lastvers() { return g_lastvers; }
void note_version(size_t v) { if (v > g_lastvers) g_lastvers = v; }

static uint64_t g_seed = 0x5eed5eedULL, g_rng_counter = 0;
static bool g_rng_dirty = true;  // host (seed, offset) not yet pushed to the device generator
void seed(uint64_t s) { g_seed = s; g_rng_counter = 0; g_rng_dirty = true; }
void rng_flush() {
  if (!g_rng_dirty) return;
  cuda::check(tcr_rand_seed(g_seed, g_rng_counter), "tcr_rand_seed");
  g_rng_dirty = false;
}
uint64_t rng_seed() { return g_seed; }
uint64_t rng_advance(uint64_t n) {
  uint64_t off = g_rng_counter;
  g_rng_counter += n;
  return off;
}

// ---------------------------------------------------------------- Variable
Variable::Variable(const void* host_data, egen::_GENERATED_DTYPE dtype, Shape shape, std::string label, Usage usage)
    : ref_(std::make_shared<cuda::DevSrc>(host_data, dtype, shape, false)), shape_(shape), label_(std::move(label)), meta_(dtype, 1), usage_(usage) {
  note_version(1);  // leaves are born at version 1 (variable.hpp:153, constant.hpp:102): a fresh functor (version 0) is stale against them
}

Variable::Variable(const Variable& other)
    : ref_(std::make_shared<cuda::DevSrc>(const_cast<Variable&>(other).ref_->data(), other.meta_.dtype_, other.shape_, false)),
      shape_(other.shape_), label_(other.label_), meta_(other.meta_), usage_(other.usage_) {}

Variable* Variable::get(const void* host_data, egen::_GENERATED_DTYPE dtype, Shape shape, std::string label, Usage usage) {
  return new Variable(host_data, dtype, shape, std::move(label), usage);
}

void Variable::upversion(size_t version) {
  meta_.version_ = std::max(meta_.version_, version);
  note_version(meta_.version_);
}

void Variable::assign(const void* input, egen::_GENERATED_DTYPE dtype, Shape shape) {
  if (false == shape.compatible_after(shape_, 0))
    global::fatalf("assigning data shaped %s to tensor %s", shape.to_string().c_str(), shape_.to_string().c_str());
  upversion(get_lastvers() + 1);
  if (dtype == meta_.dtype_) {
    ref_->assign_host(input);
    return;
  }
  size_t n = shape_.n_elems();
  std::vector<char> tmp(n * egen::type_size(meta_.dtype_));
  egen::type_convert(tmp.data(), meta_.dtype_, input, dtype, n);
  ref_->assign_host(tmp.data());
  if (ref_->resident()) cuda::sync();  // tmp goes out of scope
}

void Variable::prefetch(const void* input, egen::_GENERATED_DTYPE dtype, Shape shape) {
  if (false == shape.compatible_after(shape_, 0))
    global::fatalf("assigning data shaped %s to tensor %s", shape.to_string().c_str(), shape_.to_string().c_str());
  if (dtype != meta_.dtype_)
    global::fatalf("prefetch needs %s data (got %s): the copy is asynchronous, no conversion buffer outlives the call",
                   egen::name_type(meta_.dtype_).c_str(), egen::name_type(dtype).c_str());
  ref_->prefetch_host(input);
}

void Variable::commit() {
  upversion(get_lastvers() + 1);
  ref_->commit_prefetch();
}

void Variable::assign_device(const void* dev_input) {
  upversion(get_lastvers() + 1);
  ref_->assign_device(dev_input);
}

// ---------------------------------------------------------------- Constant
Constant::Constant(const void* host_data, egen::_GENERATED_DTYPE dtype, Shape shape)
    : ref_(std::make_shared<cuda::DevSrc>(host_data, dtype, shape, true)), shape_(shape), meta_(dtype, 1) {
  note_version(1);
  size_t n = shape.n_elems();
  std::vector<double> d(n);
  egen::type_convert(d.data(), egen::DOUBLE, host_data, dtype, n);
  scalar_ = std::all_of(d.begin(), d.end(), [&](double e) { return e == d[0]; });
  scalar_value_ = d[0];
}

Constant* Constant::get(const void* host_data, egen::_GENERATED_DTYPE dtype, Shape shape) { return new Constant(host_data, dtype, shape); }

std::string Constant::to_string() const {
  // const_encode (internal/teq/ileaf.hpp:49-80): scalar value or a bracketed prefix
  std::stringstream ss;
  if (scalar_) {
    ss << scalar_value_;
    return ss.str();
  }
  size_t n = shape_.n_elems(), shown = std::min<size_t>(n, 5);
  std::vector<double> d(shown);
  TCR_TYPE_LOOKUP(meta_.dtype_, T, {
    const T* p = (const T*)const_cast<Constant*>(this)->ref_->data();
    for (size_t i = 0; i < shown; ++i) d[i] = (double)p[i];
  });
  ss << "[";
  for (size_t i = 0; i < shown; ++i) ss << (i ? "\\" : "") << d[i];
  if (n > shown) ss << "\\...";
  ss << "]";
  return ss.str();
}

// ---------------------------------------------------------------- Functor
Functor::Functor(egen::_GENERATED_OPCODE opcode, egen::_GENERATED_DTYPE dtype, Shape shape, TensptrsT args, marsh::Maps&& attrs)
    : eigen::Observable(args, std::move(attrs)), opcode_(Opcode{egen::name_op(opcode), (size_t)opcode}), shape_(shape), args_(std::move(args)), meta_(dtype) {
  initialize();
}

Functor::Functor(const Functor& other) : eigen::Observable(other), opcode_(other.opcode_), shape_(other.shape_), args_(other.args_), meta_(other.meta_.dtype_) {
  for (auto& arg : args_)
    if (auto f = dynamic_cast<eigen::Observable*>(arg.get())) f->subscribe(this);
  initialize();
}

Functor::~Functor() {
  for (auto& child : args_)
    if (auto f = dynamic_cast<eigen::Observable*>(child.get())) f->unsubscribe(this);
}

Functor* Functor::get(egen::_GENERATED_OPCODE opcode, egen::_GENERATED_DTYPE dtype, TensptrsT children, marsh::Maps&& attrs) {
  if (children.empty()) global::fatalf("cannot perform `%s` without arguments", egen::name_op(opcode).c_str());
  ShapesT shapes;
  shapes.reserve(children.size());
  for (auto& c : children) shapes.push_back(c->shape());
  auto ctype = children.front()->get_meta().type_code();
  for (auto& c : children)
    if (ctype != c->get_meta().type_code()) global::fatal("children types are not all the same");
  Shape outshape = eigen::shape_parse(opcode, attrs, shapes);
  return new Functor(opcode, dtype, outshape, std::move(children), std::move(attrs));
}

void Functor::update_child(TensptrT arg, size_t index) {
  if (index >= args_.size())
    global::throw_errf("cannot replace argument %d when only there are only %d available", (int)index, (int)args_.size());
  uninitialize();
  if (auto f = dynamic_cast<eigen::Observable*>(args_[index].get())) {
    // another slot may still point at the same child
    bool still_used = false;
    for (size_t i = 0; i < args_.size(); ++i) still_used |= (i != index && args_[i] == args_[index]);
    if (!still_used) f->unsubscribe(this);
  }
  Shape nexshape = arg->shape(), curshape = args_[index]->shape();
  if (false == nexshape.compatible_after(curshape, 0))
    global::fatalf("cannot update child %d to argument with incompatible shape %s (requires shape %s)", (int)index,
                   nexshape.to_string().c_str(), curshape.to_string().c_str());
  auto nextype = arg->get_meta().type_label(), curtype = args_[index]->get_meta().type_label();
  if (curtype != nextype)
    global::fatalf("cannot update child %d to argument with different type %s (requires type %s)", (int)index, nextype.c_str(), curtype.c_str());
  args_[index] = arg;
  if (auto f = dynamic_cast<eigen::Observable*>(arg.get())) f->subscribe(this);
}

iDeviceRef& Functor::device() {
  if (false == has_data()) must_initialize();
  return *ref_;
}

const iDeviceRef& Functor::device() const {
  if (false == has_data()) global::fatal("cannot get device of uninitialized functor");
  return *ref_;
}

void Functor::uninitialize() {
  if (has_data()) {
    ref_ = nullptr;
    meta_.version_ = 0;
    for (auto& parent : subs_) parent->uninitialize();
  }
}

bool Functor::initialize() {
  if (std::all_of(args_.begin(), args_.end(), [](const TensptrT& child) {
        if (auto f = dynamic_cast<eigen::Observable*>(child.get())) return f->has_data();
        return true;
      }))
    cuda::typed_exec((egen::_GENERATED_OPCODE)opcode_.code_, meta_.dtype_, ref_, shape_, args_, *this);
  return has_data();
}

void Functor::must_initialize() {
  for (auto& child : args_) {
    auto f = dynamic_cast<eigen::Observable*>(child.get());
    if (nullptr != f && false == f->has_data()) f->must_initialize();
  }
  if (false == initialize()) global::fatal("failed to initialize");
}

bool Functor::prop_version(size_t max_version) {
  size_t des_version = 0;
  for (auto& child : args_) des_version = std::max(des_version, child->get_meta().state_version());
  // non-idempotent ops execute regardless of version (functor.hpp:246-269)
  size_t cur_version = meta_.version_;
  if (des_version <= cur_version && false == egen::is_idempotent((egen::_GENERATED_OPCODE)opcode_.code_)) des_version = cur_version + 1;
  des_version = std::min(des_version, max_version);
  bool propped = meta_.version_ < des_version;
  if (propped) {
    meta_.version_ = des_version;
    note_version(des_version);
  }
  return propped;
}

// ---------------------------------------------------------------- factories
TensptrT make_tfuncattr(egen::_GENERATED_DTYPE dtype, egen::_GENERATED_OPCODE opcode, TensptrsT children, marsh::Maps& attrs) {
  if (children.empty()) global::fatalf("cannot %s without arguments", egen::name_op(opcode).c_str());
  if (eigen::func_opt(opcode, dtype, attrs, children)) return children.front();  // FuncOpt: redundant -> the child itself
  if (opcode != egen::CAST) {  // TypeCaster (caster.hpp:10-44): mixed types get explicit CAST nodes
    for (auto& child : children) {
      if (child->get_meta().type_code() != (size_t)dtype) {
        marsh::Maps cattrs;
        eigen::pack_attr(cattrs, dtype);
        child = TensptrT(Functor::get(egen::CAST, dtype, {child}, std::move(cattrs)));
      }
    }
  }
  return TensptrT(Functor::get(opcode, dtype, std::move(children), std::move(attrs)));
}

TensptrT make_funcattr(egen::_GENERATED_OPCODE opcode, TensptrsT children, marsh::Maps& attrs) {
  for (auto& c : children)
    if (nullptr == c) global::fatalf("cannot %s with a null argument", egen::name_op(opcode).c_str());
  eigen::DTypesT dtypes;
  for (auto& c : children) dtypes.push_back((egen::_GENERATED_DTYPE)c->get_meta().type_code());
  auto typecode = eigen::type_parse(opcode, attrs, dtypes);
  return make_tfuncattr(typecode, opcode, std::move(children), attrs);
}

static std::vector<char> fill_scalar(double scalar, size_t n, egen::_GENERATED_DTYPE dtype) {
  std::vector<char> buf(n * egen::type_size(dtype));
  TCR_TYPE_LOOKUP(dtype, T, {
    T* p = (T*)buf.data();
    std::fill(p, p + n, (T)scalar);
  });
  return buf;
}

VarptrT make_variable_scalar(double scalar, Shape shape, std::string label, egen::_GENERATED_DTYPE dtype) {
  if (label.empty()) {
    std::stringstream ss;
    ss << scalar;
    label = ss.str();
  }
  auto buf = fill_scalar(scalar, shape.n_elems(), dtype);
  return VarptrT(Variable::get(buf.data(), dtype, shape, label));
}

VarptrT make_variable(const void* data, egen::_GENERATED_DTYPE dtype, Shape shape, std::string label) {
  return VarptrT(Variable::get(data, dtype, shape, std::move(label)));
}

TensptrT make_constant_tensor(const void* data, egen::_GENERATED_DTYPE dtype, Shape shape) { return TensptrT(Constant::get(data, dtype, shape)); }

TensptrT make_constant_scalar(double scalar, Shape shape, egen::_GENERATED_DTYPE dtype) {
  auto buf = fill_scalar(scalar, shape.n_elems(), dtype);
  return make_constant_tensor(buf.data(), dtype, shape);
}

TensptrT make_constant_like(double scalar, TensptrT like) {
  auto like_type = (egen::_GENERATED_DTYPE)like->get_meta().type_code();
  TensptrT cst = make_constant_scalar(scalar, Shape(), like_type);
  return make_functor(egen::EXTEND, TensptrsT{cst}, like);
}

void run(const TensptrsT& targets, const TensSetT& ignored, size_t max_version) {
  cuda::ensure_device();
  cuda::Device device(max_version);
  TensSetT targset;
  for (auto& t : targets) targset.emplace(t.get());
  teq::get_eval().evaluate(device, targset, ignored);
}

TensptrsT derive(TensptrT root, const TensptrsT& targets) {
  DerivativeFuncs builder;
  return teq::derive(root, targets, builder);
}

}  // namespace eteq
