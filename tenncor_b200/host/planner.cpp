// planner.cpp — lowers a functor DAG to a fused launch plan and replays it as a CUDA graph.
// See planner.hpp for the design; reference hot loop being replaced:
// teq::TravEvaluator::visit_func (internal/teq/evaluator.hpp:34-43) + eigen::Device::calc
// (internal/eigen/device.hpp:555-570) + TensOp::assign (device.hpp:304-328).
#include "planner.hpp"
#include "dp.hpp"
#include <queue>
#include <set>
#include <map>
#include <unordered_map>
#include <limits>
#include <tuple>
#include <algorithm>

#include <cstdlib>
#include <cstring>

namespace cuda {

using namespace teq;
using namespace egen;

static PlanStats g_stats;
PlanStats last_plan_stats() { return g_stats; }
struct Plan;
static Plan* g_last_plan = nullptr;

namespace {

constexpr int MAX_IN = TCR_EW_MAX_INPUTS, MAX_INSTR = TCR_EW_MAX_INSTRS, NREGS = TCR_EW_NREGS;

bool is_unary(int op) { return op >= ABS && op <= CUBE; }
bool is_binary(int op) { return op == POW || op == SUB || op == DIV || op == MIN || op == MAX || op == EQ || op == NEQ || op == LT || op == GT; }
bool is_assign(int op) { return op >= ASSIGN && op <= ASSIGN_DIV; }
bool is_ew_op(int op) { return is_unary(op) || is_binary(op) || op == ADD || op == MUL || op == SELECT || op == CAST || is_assign(op); }
bool has_compute_kernels(_GENERATED_DTYPE t) { return t == FLOAT || t == DOUBLE || t == INT32 || t == INT64; }

struct PNode {
  iTensor* tens = nullptr;
  iFunctor* func = nullptr;  // null: leaf (or ignored functor treated as a leaf)
  eigen::iEigen* holder = nullptr;
  int op = 0;
  _GENERATED_DTYPE dtype = BAD_TYPE;
  Shape shape;
  int64_t n = 1;
  std::vector<int> args;
  int root = -1;          // storage base node (self unless a view / assign alias)
  size_t offset = 0;      // byte offset into the base node's buffer
  std::vector<int> consumers;
  bool exposed = false, inlined = false, is_view = false, is_ew = false, is_extend = false, needs_mat = false;
  bool has_scalar = false;  // value is one scalar known at plan time (constant leaf / EXTEND of one)
  double scalar = 0;
  bool mutable_leaf = false;
  void* ptr = nullptr;
  int step = -1;
  int region = -1;  // root node of the EW region this node is computed in
  // GEMM with fused epilogue: this node (ADD of a bias, optionally through SIGMOID / TANH) is
  // produced by the GEMM kernel of node `gemm_src`
  int gemm_src = -1, gemm_bias = -1, gemm_epi = 0, gemm_act = 0;
  bool gemm_transposed = false;  // PERMUTE{1,0} of a GEMM: computed as (B^T A^T), no copy
  int64_t bucket_slot = -1;      // >= 0: this node's buffer is the gradient bucket at this byte offset
  int conv = -1;                 // >= 0: produced by the patch-gather + GEMM step `convs[conv]` (fuse_convs)
  int stack = -1, stack_pos = 0;  // >= 0: the buffer is slab `stack_pos` of the plan-owned stack `stack` (fuse_recurrent_steps)
  int alias_stack = -1;           // >= 0: the node IS that whole stack (an n-ary CONCAT whose operands were produced in place)
};

struct InputRef {
  int node;       // base node whose buffer is read
  size_t offset;  // byte offset
  uint32_t mask;  // ranks (of the region root's shape) along which the input is broadcast
  _GENERATED_DTYPE dtype;
  bool operator==(const InputRef& o) const { return node == o.node && offset == o.offset && mask == o.mask; }
};

struct Expr {
  int kind;  // 0 input, 1 const, 2 op
  int op = 0, a = -1, b = -1, c = -1, input = -1;
  double imm = 0;
  int uses = 0, reg = -1;
  bool emitted = false;
};

struct Step {
  bool ew = false;
  // ew
  tcr_ew_program prog;
  std::vector<InputRef> inputs;
  int out_node = -1;
  // launch
  const DevOp* holder = nullptr;
  void* out = nullptr;
  std::vector<const void*> in;
  std::vector<int> in_nodes;
  std::vector<size_t> in_offsets;
  // fused GEMM epilogue (in[2] = bias)
  bool gemm_fused = false;
  tcr_gemm_desc gemm;
  // conv2d composite as patch gather + GEMM (fuse_convs): in = {image, kernel | upstream gradient, [bias]}
  bool conv_fused = false, conv_grad = false, conv_dimg = false;
  size_t conv_out_offset = 0;
  int64_t conv_img[8], conv_win[8], conv_pitch = 0, conv_rows = 0;
  void* conv_cols = nullptr;
  // data-parallel gradient bucket: member = a gradient placed in the bucket (copy, or nothing when
  // its producer already wrote there); flush = ONE all-reduce over the whole bucket
  enum Kind { NORMAL = 0, BUCKET_MEMBER, BUCKET_FLUSH } kind = NORMAL;
  bool member_in_place = false;
  size_t bucket_offset = 0;         // member: byte offset of its slot
  std::vector<int> bucket_members;  // flush: member step indices
  // ---- recurrent-step fusion (fuse_recurrent_steps)
  // grouped / K-segmented small-batch product (tcr_gemm_grouped): in = {A segments, B[group][segment], biases present, c_{t-1}}
  bool group = false;
  tcr_gemm_group_desc gd;
  int g_bias_slot[4] = {-1, -1, -1, -1};  // index into `in` of the bias of group g
  int g_cprev_slot = -1;
  int g_out_nodes[4] = {-1, -1, -1, -1};  // node each group's result is stored to (-1: not stored)
  int g_c_node = -1, g_h_node = -1;       // LSTM cell epilogue outputs
  std::vector<int> extra_outs;            // nodes written besides out_node
  std::vector<int> dep_nodes;             // nodes read through a stack base pointer: ordering / liveness only
  // REDUCE over a stack of operands laid out back to back (the 128 per-step bias gradients as one reduction)
  bool stack_reduce = false;
  int red_op = 0;
  uint32_t red_mask = 0;
  int64_t red_shape[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  int64_t pos = 0;  // ordering key of the stable topological re-sort
  bool dead = false;
  int gemm_aux_slot = -1;  // index into `in` of the operand of the product's post-op (activation derivative)
  // map-reduce (tcr_elementwise_reduce): the program's output is summed over every element into `out`, then scaled
  bool ew_reduce = false;
  int red_post = 0;
  double red_imm = 0;
  // several elementwise programs over one iteration space in one launch (tcr_elementwise_multi, merge_elementwise_steps): program k
  // reads multi_inputs[k] and stores node multi_outs[k]; `inputs` holds the reads that come from outside the step
  bool multi = false;
  std::vector<tcr_ew_program> multi_progs;
  std::vector<std::vector<InputRef>> multi_inputs;
  std::vector<int> multi_outs;
  std::vector<uint8_t> multi_keep;  // 0: the result is read only by later programs of the launch and is not written to memory
  // the multi launch has the shape of a gated cell's backward step (match_cell_backward): program 0 = s, 1 = c, then the gates
  // several CONCAT steps as one batched 2-D copy launch (batch_concat_steps): item k copies into node copy_dst[k] at byte copy_off[k]
  bool copy_batch = false;
  std::vector<tcr_copy2d_item> copy_items;
  std::vector<int> copy_dst, copy_src;
  std::vector<size_t> copy_dst_off, copy_src_off;
  bool cell_bwd = false;
  int cb_c[3][2] = {{0, 0}, {0, 0}, {0, 0}};  // (program, input) of c_x, c_y, c_z
  int cb_kind[6] = {0, 0, 0, 0, 0, 0}, cb_sel[6] = {0, 0, 0, 0, 0, 0}, cb_x[6] = {0, 0, 0, 0, 0, 0}, cb_y[6] = {0, 0, 0, 0, 0, 0};  // gate k = program k + 2; x / y input slots
};

// One recognised conv2d composite (cfg/tenncor/nn.yml:48-98) or its kernel gradient
// (tenncor/eteq/backprop.hpp CONV rule, arg 1): see Plan::fuse_convs.
struct ConvFuse {
  bool grad = false;
  bool dimg = false;         // image gradient: GEMM into the patch matrix, then col2im into one slice of the CONV's output
  size_t out_offset = 0;     // dimg: byte offset of that slice (the only part of the output any reader touches)
  int64_t out_n = 0;         // dimg: elements of the slice
  int img = -1, other = -1;  // arg nodes read directly: the un-padded image; the kernel (forward) / upstream gradient (grad)
  int first = 0;             // earliest node position whose read moves to the fused step
  int64_t img_shape[8], win[8], rows = 1, k = 1, pitch = 0;
  tcr_gemm_desc gemm;
};

}  // namespace

struct Plan {
  std::vector<PNode> nodes;
  std::vector<Step> steps;
  std::vector<std::weak_ptr<iTensor>> target_refs;
  std::vector<iTensor*> target_keys;
  std::vector<int> functor_order;  // indices of functor nodes in evaluation order
  std::vector<std::pair<void*, size_t>> owned;  // plan-owned device buffers
  std::vector<std::pair<int, void*>> bound;     // (node, pointer) of holder / leaf buffers baked into the plan
  eigen::RTMemptrT memory;
  void* graph = nullptr;
  bool use_graph = false, has_run = false, always_run = false;
  int precision = 0;
  size_t n_launch_steps = 0;
  // gradient bucket (data parallel): one flat buffer, one collective per step
  bool uses_rand = false;
  void* bucket = nullptr;
  size_t bucket_bytes = 0;
  int bucket_dtype = 0;
  double bucket_scale = 1.0;

  ~Plan() {
    if (graph) tcr_graph_destroy(graph);
    for (auto& b : owned) memory->deallocate(b.first, b.second);
  }

  // ---------------------------------------------------------------- graph collection
  std::unordered_map<iTensor*, int> index;

  int collect(iTensor* t, const TensSetT& ignored) {
    auto it = index.find(t);
    if (it != index.end()) return it->second;
    PNode n;
    n.tens = t;
    n.dtype = (_GENERATED_DTYPE)t->get_meta().type_code();
    n.shape = t->shape();
    n.n = (int64_t)n.shape.n_elems();
    auto f = dynamic_cast<iFunctor*>(t);
    if (f && !ignored.count(t)) {
      for (auto& a : f->args_ref()) n.args.push_back(collect(a.get(), ignored));
      n.func = f;
      n.op = (int)f->get_opcode().code_;
      n.holder = &static_cast<eigen::iEigen&>(f->device());
    } else {
      if (auto c = dynamic_cast<eteq::Constant*>(t)) {
        n.has_scalar = c->is_scalar();
        n.scalar = c->scalar_value();
      }
      n.mutable_leaf = nullptr != dynamic_cast<eigen::iMutableLeaf*>(t);
    }
    int id = (int)nodes.size();
    n.root = id;
    nodes.push_back(std::move(n));
    index.emplace(t, id);
    if (nodes[id].func) functor_order.push_back(id);
    return id;
  }

  std::pair<int, size_t> base(int i) const { return {nodes[i].root, nodes[i].offset}; }

  void resolve_views() {
    for (size_t i = 0; i < nodes.size(); ++i) {
      PNode& n = nodes[i];
      if (!n.func) continue;
      if (auto ref = dynamic_cast<DevRef*>(n.holder)) {
        int a = index.at(ref->referent());
        n.is_view = true;
        n.root = nodes[a].root;
        n.offset = nodes[a].offset + ref->byte_offset();
        if (nodes[a].has_scalar && ref->byte_offset() == 0) { n.has_scalar = true; n.scalar = nodes[a].scalar; }
      } else if (dynamic_cast<DevAssign*>(n.holder)) {
        n.is_ew = true;  // computed by an EW program writing the variable's storage; consumers read the variable
      } else if (n.op == EXTEND) {
        n.is_extend = true;
        const PNode& src = nodes[n.args[0]];
        if (src.has_scalar) { n.has_scalar = true; n.scalar = src.scalar; }
      } else if (is_ew_op(n.op) && has_compute_kernels(n.dtype)) {
        n.is_ew = true;
        // a CAST from a storage-only type (UINT8 pixels, INT16 ...) into a compute type converts at load inside the program that
        // consumes it (tenncor/eteq/caster.hpp:10-44 static_casts element-wise, as load_any does): no separate CAST launch
      }
    }
  }

  // ---------------------------------------------------------------- EW regions
  std::unordered_map<int, std::vector<int>> members;       // region root -> member nodes (incl. root)
  std::unordered_map<int, std::vector<int>> assign_pos;    // mutable leaf -> positions of ASSIGN nodes writing it
  std::vector<ConvFuse> convs;

  uint32_t bcast_mask(const Shape& src, const Shape& dst) const {
    uint32_t m = 0;
    for (int r = 0; r < rank_cap; ++r)
      if (dst.at(r) > 1 && src.at(r) == 1) m |= 1u << r;
    return m;
  }

  struct Gen {
    std::vector<Expr> ex;
    std::vector<InputRef> inputs;
    std::unordered_map<int, int> memo;  // member node -> expr
    bool ok = true;
  };

  int add_input(Gen& g, InputRef ref) {
    for (size_t k = 0; k < g.inputs.size(); ++k)
      if (g.inputs[k] == ref) { Expr e; e.kind = 0; e.input = (int)k; g.ex.push_back(e); return (int)g.ex.size() - 1; }
    if ((int)g.inputs.size() >= MAX_IN) { g.ok = false; return -1; }
    g.inputs.push_back(ref);
    Expr e;
    e.kind = 0;
    e.input = (int)g.inputs.size() - 1;
    g.ex.push_back(e);
    return (int)g.ex.size() - 1;
  }

  int add_op(Gen& g, int op, int a, int b = -1, int c = -1) {
    Expr e;
    e.kind = 2; e.op = op; e.a = a; e.b = b; e.c = c;
    g.ex.push_back(e);
    return (int)g.ex.size() - 1;
  }

  // expression of node `i` as seen by region `root`
  int build_expr(Gen& g, int root, int i) {
    if (!g.ok) return -1;
    const PNode& root_n = nodes[root];
    const PNode& ni = nodes[i];
    if (ni.has_scalar && !(i == root)) {
      Expr e;
      e.kind = 1;
      e.imm = ni.scalar;
      g.ex.push_back(e);
      return (int)g.ex.size() - 1;
    }
    int r = ni.root;
    size_t off = ni.offset;
    const PNode& nr = nodes[r];
    bool member = (r == root && i == root) || (off == 0 && nr.region == root && (nr.inlined || r == root) && !ni.is_view) ||
                  (off == 0 && ni.is_view && nr.region == root && (nr.inlined));
    if (r == root && i != root) member = false;  // a consumer inside the region cannot be upstream of the root
    if (member) {
      auto it = g.memo.find(r);
      if (it != g.memo.end()) return it->second;
      int e = -1;
      int op = nr.op;
      if (is_assign(op)) {
        int var = build_input(g, root, nr.args[0]);
        int src = build_expr(g, root, nr.args[1]);
        if (op == ASSIGN) e = add_op(g, TCR_EW_MOV, src);
        else e = add_op(g, op == ASSIGN_ADD ? ADD : op == ASSIGN_SUB ? SUB : op == ASSIGN_MUL ? MUL : DIV, var, src);
      } else if (op == CAST) {
        e = add_op(g, TCR_EW_MOV, build_input(g, root, nr.args[0]));  // conversion happens at load
      } else if (op == ADD || op == MUL) {
        e = build_expr(g, root, nr.args[0]);
        for (size_t k = 1; k < nr.args.size() && g.ok; ++k) e = add_op(g, op, e, build_expr(g, root, nr.args[k]));
      } else if (op == SELECT) {
        int a = build_expr(g, root, nr.args[0]), b = build_expr(g, root, nr.args[1]), c = build_expr(g, root, nr.args[2]);
        e = add_op(g, TCR_EW_SELECT, a, b, c);
      } else if (is_unary(op)) {
        e = add_op(g, op, build_expr(g, root, nr.args[0]));
      } else {
        int a = build_expr(g, root, nr.args[0]), b = build_expr(g, root, nr.args[1]);
        e = add_op(g, op, a, b);
      }
      g.memo.emplace(r, e);
      (void)root_n;
      return e;
    }
    return build_input(g, root, i);
  }

  int build_input(Gen& g, int root, int i) {
    const PNode& ni = nodes[i];
    const PNode& root_n = nodes[root];
    // broadcast read through an EXTEND that nobody needs materialised
    if (ni.is_extend && !ni.needs_mat && ni.shape == root_n.shape) {
      const PNode& src = nodes[ni.args[0]];
      return add_input(g, InputRef{src.root, src.offset, bcast_mask(src.shape, root_n.shape), src.dtype});
    }
    if (ni.n != root_n.n) { g.ok = false; return -1; }
    return add_input(g, InputRef{ni.root, ni.offset, 0, ni.dtype});
  }

  // register allocation + instruction emission; fills `step`
  bool emit(Gen& g, int root, int root_expr, Step& step) {
    if (!g.ok) return false;
    const PNode& rn = nodes[root];
    // iteration-space split shared by every broadcast input (<= 3 segments)
    int64_t dims[3] = {1, 1, 1};
    int seg_of_rank[rank_cap];
    {
      int seg = -1;
      uint32_t prev_sig = 0;
      bool first = true;
      for (int r = 0; r < rank_cap; ++r) {
        seg_of_rank[r] = seg < 0 ? 0 : seg;
        if (rn.shape.at(r) == 1) continue;
        uint32_t sig = 0;
        for (size_t k = 0; k < g.inputs.size(); ++k) sig |= ((g.inputs[k].mask >> r) & 1u) << k;
        if (first || sig != prev_sig) { ++seg; first = false; prev_sig = sig; }
        if (seg >= 3) return false;
        seg_of_rank[r] = seg;
        dims[seg] *= rn.shape.at(r);
      }
    }
    // use counts
    std::vector<int> order;
    std::function<void(int)> count = [&](int e) {
      Expr& x = g.ex[e];
      x.uses++;
      if (x.uses > 1) return;
      if (x.kind == 2) { count(x.a); if (x.b >= 0) count(x.b); if (x.c >= 0) count(x.c); }
      order.push_back(e);
    };
    count(root_expr);
    int owner[NREGS];
    for (int r = 0; r < NREGS; ++r) owner[r] = -1;
    // an input expr may appear several times (deduped inputs share a register)
    std::vector<int> input_uses(g.inputs.size(), 0);
    for (int e : order)
      if (g.ex[e].kind == 0) input_uses[g.ex[e].input] += g.ex[e].uses;
    for (size_t k = 0; k < g.inputs.size(); ++k) owner[k] = 1000 + (int)k;  // held by input k
    std::memset(&step.prog, 0, sizeof(step.prog));
    tcr_ew_program& p = step.prog;
    int ninstr = 0;
    auto release = [&](int e) {
      Expr& x = g.ex[e];
      if (x.kind == 0) {
        if (--input_uses[x.input] == 0) owner[x.input] = -1;
      } else if (--x.uses == 0) {
        owner[x.reg] = -1;
      }
    };
    auto reg_of = [&](int e) { return g.ex[e].kind == 0 ? g.ex[e].input : g.ex[e].reg; };
    for (int e : order) {
      Expr& x = g.ex[e];
      if (x.kind == 0) continue;
      int ra = 0, rb = 0, rc = 0;
      if (x.kind == 2) {
        ra = reg_of(x.a);
        if (x.b >= 0) rb = reg_of(x.b);
        if (x.c >= 0) rc = reg_of(x.c);
        release(x.a);
        if (x.b >= 0) release(x.b);
        if (x.c >= 0) release(x.c);
      }
      int dst = -1;
      for (int r = 0; r < NREGS; ++r)
        if (owner[r] == -1) { dst = r; break; }
      if (dst < 0 || ninstr >= MAX_INSTR) return false;
      owner[dst] = e;
      x.reg = dst;
      tcr_ew_instr& ins = p.instrs[ninstr++];
      ins.op = (uint8_t)(x.kind == 1 ? TCR_EW_CONST : x.op);
      ins.dst = (uint8_t)dst;
      ins.a = (uint8_t)ra; ins.b = (uint8_t)rb; ins.c = (uint8_t)rc;
      ins.imm = x.imm;
    }
    // a CAST root converts at load (from any storage type): compute in the OUTPUT type so that the store is a plain copy
    p.dtype = rn.dtype;
    if (!has_compute_kernels((_GENERATED_DTYPE)p.dtype)) return false;
    p.n_inputs = (int)g.inputs.size();
    p.n_outputs = 1;
    p.n_instrs = ninstr;
    for (int k = 0; k < 3; ++k) p.dims[k] = dims[k];
    for (size_t k = 0; k < g.inputs.size(); ++k) {
      p.inputs[k].dtype = g.inputs[k].dtype;
      for (int r = 0; r < rank_cap; ++r)
        if (rn.shape.at(r) > 1 && ((g.inputs[k].mask >> r) & 1u)) p.inputs[k].bcast[seg_of_rank[r]] = 1;
    }
    p.outputs[0].dtype = rn.dtype;
    p.outputs[0].reg = (uint8_t)reg_of(root_expr);
    step.inputs = g.inputs;
    step.out_node = root;
    step.ew = true;
    return true;
  }

  bool try_region(int root, Step& step) {
    Gen g;
    int e = build_expr(g, root, root);
    return emit(g, root, e, step);
  }

  // would moving region `a` (first evaluated at first_pos) to position `to` cross an ASSIGN of a leaf it reads?
  bool crosses_assign(int a_root, int to) {
    int first_pos = a_root;
    for (int m : members[a_root]) first_pos = std::min(first_pos, m);
    for (int m : members[a_root]) {
      for (int arg : nodes[m].args) {
        int leaf = nodes[arg].root;
        if (nodes[arg].is_extend) leaf = nodes[nodes[arg].args[0]].root;
        auto it = assign_pos.find(leaf);
        if (it == assign_pos.end()) continue;
        for (int pos : it->second)
          if (pos > first_pos && pos < to) return true;
      }
    }
    return false;
  }


  // ---------------------------------------------------------------- conv2d composite -> patch gather + GEMM
  // The reference has no multi-filter convolution opcode. nn.conv2d (cfg/tenncor/nn.yml:48-98) is
  //   PERMUTE(CONV(PAD(img, fresh rank d by (out-1, out-1)), REVERSE(kernel, {0}), {d, 0, 1, 2}), {d, 1, 2, 3})
  // i.e. the single-kernel N-d valid correlation (operator.hpp:1143-1187) slides the reversed kernel along a
  // zero-padded rank so that each of its `out` positions meets exactly one filter. Seen whole it is
  //   out[(x,y,b), o] = sum_{c,i,j} img[c, x+i, y+j, b] * kernel[o, c, i, j]
  // = cols[(x,y,b), (c,i,j)] . kernel[(c,i,j), o]: a patch gather (tcr_im2col) and ONE tensor-core GEMM whose
  // row-major output already has the PERMUTE's layout. Its kernel gradient (backprop.hpp CONV rule, arg 1, followed
  // by the PERMUTE / REVERSE rules) REVERSE(PERMUTE(CONV(PAD(img), sup', identity))) is cols^T . sup in the kernel's
  // own layout. Anything that does not match keeps the generic CONV kernel.
  static RanksT full_order(const iFunctor& f) {
    RanksT order = eigen::unpack_ranks(f);
    if (order.size() > rank_cap) order.resize(rank_cap);
    bool visited[rank_cap] = {false};
    for (auto r : order) {
      if (r >= rank_cap || visited[r]) return {};
      visited[r] = true;
    }
    for (RankT i = 0; i < rank_cap; ++i)
      if (!visited[i]) order.push_back(i);
    return order;
  }

  // the node whose buffer `arg` names element for element (looks through IDENTITY-like views), or -1
  int solid(int arg) const {
    const PNode& a = nodes[arg];
    if (a.offset != 0 || !(nodes[a.root].shape == a.shape)) return -1;
    return a.root;
  }

  struct PadTrick {
    int pad = -1, img = -1, d = -1;
    int64_t p = 0;
  };

  // PAD of one singular trailing rank d by (p, p), p >= 1
  bool pad_trick(int arg, PadTrick& t) const {
    const int xi = solid(arg);
    if (xi < 0) return false;
    const PNode& x = nodes[xi];
    if (!x.func || x.op != PAD || x.dtype != FLOAT || x.args.size() != 1) return false;
    auto pads = eigen::unpack_dimpairs(*x.func);
    int d = -1;
    int64_t p = 0;
    for (size_t r = 0; r < pads.size() && r < rank_cap; ++r) {
      if (pads[r].first == 0 && pads[r].second == 0) continue;
      if (d >= 0 || pads[r].first != pads[r].second) return false;
      d = (int)r;
      p = pads[r].first;
    }
    if (d < 0 || p < 1) return false;
    const PNode& img = nodes[x.args[0]];
    if (img.dtype != FLOAT) return false;
    for (int r = d; r < rank_cap; ++r)
      if (img.shape.at(r) != 1) return false;
    t.pad = xi; t.img = x.args[0]; t.d = d; t.p = p;
    return true;
  }

  // in the memory order of PERMUTE(src, perm): is d the fastest non-singular rank and are the others ascending?
  static bool d_first(const Shape& src, const RanksT& perm, int d) {
    int prev = -1;
    bool first = true;
    for (int q = 0; q < rank_cap; ++q) {
      const int r = perm[q];
      if (src.at(r) == 1) continue;
      if (first) {
        if (r != d) return false;
        first = false;
        continue;
      }
      if (r <= prev) return false;
      prev = r;
    }
    return !first;
  }

  bool assigned_between(std::initializer_list<int> arg_nodes, int lo, int hi) const {
    for (int a : arg_nodes) {
      auto it = assign_pos.find(nodes[a].root);
      if (it == assign_pos.end()) continue;
      for (int pos : it->second)
        if (pos > lo && pos < hi) return true;
    }
    return false;
  }

  void fuse_convs() {
    if (std::getenv("TCR_NO_CONV_GEMM")) return;
    std::unordered_map<int, size_t> skipped;  // helper node (PAD / REVERSE / PERMUTE) -> consumers that now read through it
    std::vector<std::pair<int, int>> chains;  // (helper, its helper argument): once the first is inlined the second lost a consumer
    for (size_t i = 0; i < nodes.size(); ++i) {
      PNode& n = nodes[i];
      if (!n.func || n.is_view || n.dtype != FLOAT) continue;
      if (n.op == CONV && n.args.size() == 2) {
        fuse_conv_image_gradient((int)i, skipped, chains);
        continue;
      }
      if (n.args.size() != 1) continue;
      if (n.op == PERMUTE) {
        // ---- forward
        const int ci = solid(n.args[0]);
        if (ci < 0) continue;
        PNode& c = nodes[ci];
        if (!c.func || c.op != CONV || c.dtype != FLOAT || c.exposed || c.inlined || c.consumers.size() != 1 || c.args.size() != 2) continue;
        const RanksT order = full_order(*c.func), perm = full_order(*n.func);
        if (order.empty() || perm.empty()) continue;
        PadTrick t;
        if (!pad_trick(c.args[0], t) || order[0] != t.d) continue;
        const int ki = solid(c.args[1]);
        if (ki < 0) continue;
        const PNode& kr = nodes[ki];
        if (!kr.func || kr.op != REVERSE || kr.dtype != FLOAT || kr.args.size() != 1) continue;
        auto rset = eigen::unpack_rankset(*kr.func);
        if (rset.size() != 1 || *rset.begin() != 0 || (int64_t)kr.shape.at(0) != t.p + 1) continue;
        ConvFuse cf;
        const PNode& img = nodes[t.img];
        for (int r = 0; r < rank_cap; ++r) { cf.img_shape[r] = img.shape.at(r); cf.win[r] = 1; }
        bool ok = true;
        int prev = -1;
        for (int q = 1; q < rank_cap; ++q) {  // kernel rank q slides along image rank order[q]
          if (kr.shape.at(q) == 1) continue;
          const int r = order[q];
          if (r <= prev || r >= t.d) ok = false;  // window coordinates must enumerate in the kernel's memory order
          prev = r;
          if (ok) cf.win[r] = kr.shape.at(q);
        }
        for (int r = 0; r < rank_cap && ok; ++r) {
          const int64_t e = r == t.d ? t.p + 1 : cf.img_shape[r] - cf.win[r] + 1;
          if (e < 1 || (int64_t)c.shape.at(r) != e) ok = false;
          if (r != t.d) { cf.rows *= e; cf.k *= cf.win[r]; }
        }
        if (!ok || !d_first(c.shape, perm, t.d)) continue;
        cf.pitch = (cf.k + 3) / 4 * 4;
        cf.img = t.img;
        cf.other = kr.args[0];
        cf.first = std::min({t.pad, ki, ci});
        if (assigned_between({cf.img, cf.other}, cf.first, (int)i)) continue;
        tcr_gemm_desc& g = cf.gemm;
        std::memset(&g, 0, sizeof(g));
        g.m = cf.rows; g.n = t.p + 1; g.k = cf.k; g.batch = 1;
        g.a_sm = cf.pitch; g.a_sk = 1;
        g.b_sk = g.n; g.b_sn = 1;
        g.c_sm = g.n; g.c_sn = 1;
        g.dtype = FLOAT;
        n.conv = (int)convs.size();
        convs.push_back(cf);
        c.inlined = true;
        ++skipped[t.pad];
        ++skipped[ki];
      } else if (n.op == REVERSE) {
        // ---- kernel gradient
        auto rset = eigen::unpack_rankset(*n.func);
        if (rset.size() != 1 || *rset.begin() != 0) continue;
        const int pi = solid(n.args[0]);
        if (pi < 0) continue;
        PNode& p2 = nodes[pi];
        if (!p2.func || p2.op != PERMUTE || p2.dtype != FLOAT || p2.exposed || p2.inlined || p2.conv >= 0 || p2.consumers.size() != 1 || p2.args.size() != 1) continue;
        const int ci = solid(p2.args[0]);
        if (ci < 0) continue;
        PNode& c = nodes[ci];
        if (!c.func || c.op != CONV || c.dtype != FLOAT || c.exposed || c.inlined || c.consumers.size() != 1 || c.args.size() != 2) continue;
        const RanksT order = full_order(*c.func), perm = full_order(*p2.func);
        if (order.empty() || perm.empty()) continue;
        PadTrick t;
        if (!pad_trick(c.args[0], t)) continue;
        const Shape ss = nodes[c.args[1]].shape;  // the upstream gradient in the CONV's output layout
        if (nodes[c.args[1]].dtype != FLOAT) continue;
        bool ok = (int64_t)ss.at(t.d) == t.p + 1 && (int64_t)n.shape.at(0) == t.p + 1;
        for (int q = 0; q < rank_cap; ++q)
          if (ss.at(q) > 1 && order[q] != q) ok = false;
        ConvFuse cf;
        cf.grad = true;
        const PNode& img = nodes[t.img];
        for (int r = 0; r < rank_cap && ok; ++r) {
          cf.img_shape[r] = img.shape.at(r);
          cf.win[r] = 1;
          if (r == t.d) { ok = (int64_t)c.shape.at(r) == t.p + 1; continue; }
          cf.win[r] = cf.img_shape[r] - (int64_t)ss.at(r) + 1;
          if (cf.win[r] < 1 || (int64_t)c.shape.at(r) != cf.win[r]) ok = false;
          cf.rows *= ss.at(r);
          cf.k *= cf.win[r];
        }
        if (!ok || !d_first(c.shape, perm, t.d)) continue;
        cf.pitch = (cf.k + 3) / 4 * 4;
        cf.img = t.img;
        cf.other = c.args[1];
        cf.first = std::min({t.pad, ci, pi});
        tcr_gemm_desc& g = cf.gemm;
        std::memset(&g, 0, sizeof(g));
        g.m = cf.k; g.n = t.p + 1; g.k = cf.rows; g.batch = 1;
        g.a_sm = 1; g.a_sk = cf.pitch;      // cols^T
        g.b_sk = 1; g.b_sn = cf.rows;       // sup' [positions..., out]: out slowest
        g.c_sm = g.n; g.c_sn = 1;
        g.dtype = FLOAT;
        // sup' is normally PERMUTE(sup) of a gradient laid out [out, positions...] (the PERMUTE rule): read that instead
        int through = -1;
        const int si = solid(c.args[1]);
        if (si >= 0 && nodes[si].func && nodes[si].op == PERMUTE && nodes[si].args.size() == 1 && nodes[nodes[si].args[0]].dtype == FLOAT) {
          const RanksT sperm = full_order(*nodes[si].func);
          if (!sperm.empty()) {
            RanksT inv(rank_cap);
            for (int q = 0; q < rank_cap; ++q) inv[sperm[q]] = (RankT)q;
            if (d_first(ss, inv, t.d)) through = si;
          }
        }
        if (through >= 0) {
          cf.other = nodes[through].args[0];
          cf.first = std::min(cf.first, through);
          g.b_sk = g.n; g.b_sn = 1;
        }
        if (assigned_between({cf.img, cf.other}, cf.first, (int)i)) continue;
        n.conv = (int)convs.size();
        convs.push_back(cf);
        c.inlined = true;
        p2.inlined = true;
        ++skipped[t.pad];
        if (through >= 0) ++skipped[through];
      }
    }
    std::vector<char> chain_done(chains.size(), 0);
    for (bool changed = true; changed;) {
      changed = false;
      for (auto& kv : skipped) {
        PNode& x = nodes[kv.first];
        if (!x.exposed && !x.inlined && kv.second == x.consumers.size()) { x.inlined = true; changed = true; }
      }
      for (size_t c = 0; c < chains.size(); ++c)
        if (!chain_done[c] && nodes[chains[c].first].inlined) { ++skipped[chains[c].second]; chain_done[c] = 1; changed = true; }
    }
  }

  // Image gradient of the conv2d composite (backprop.hpp CONV rule, arg 0): CONV(PAD(sup', kernel extents - 1 on every
  // slid rank), REVERSE(REVERSE(kernel, {0}), all ranks), same order), of which the PAD rule's SLICE keeps only position
  // p of rank d. That slice is dimg[c,x,y,b] = sum_{o,i,j} sup[o, x-i, y-j, b] * kernel[o,c,i,j]: one GEMM
  // cols[(x',y',b), (c,i,j)] = sup[(x',y',b), o] . kernel[o, (c,i,j)] and the adjoint of the patch gather (tcr_col2im).
  // The other 2p positions of rank d — (2*out - 1)/1 of the reference's work — are never read and never computed.
  void fuse_conv_image_gradient(int i, std::unordered_map<int, size_t>& skipped, std::vector<std::pair<int, int>>& chains) {
    PNode& n = nodes[i];
    if (n.inlined || n.conv >= 0) return;
    const RanksT order = full_order(*n.func);
    if (order.empty()) return;
    const int d = order[0];
    // kernel operand: REVERSE over every non-singular rank of REVERSE(kernel, {0})
    const int ra = solid(n.args[1]);
    if (ra < 0) return;
    const PNode& rev_all = nodes[ra];
    if (!rev_all.func || rev_all.op != REVERSE || rev_all.dtype != FLOAT || rev_all.args.size() != 1) return;
    const Shape ks = rev_all.shape;
    auto rs = eigen::unpack_rankset(*rev_all.func);
    for (int q = 0; q < rank_cap; ++q)
      if (ks.at(q) > 1 && !rs.count((RankT)q)) return;
    const int kr = solid(rev_all.args[0]);
    if (kr < 0) return;
    const PNode& rev0 = nodes[kr];
    if (!rev0.func || rev0.op != REVERSE || rev0.dtype != FLOAT || rev0.args.size() != 1) return;
    auto rs0 = eigen::unpack_rankset(*rev0.func);
    if (rs0.size() != 1 || *rs0.begin() != 0) return;
    const int64_t nout = ks.at(0), p = nout - 1;
    if (p < 1) return;
    // image operand: PAD of the upstream gradient by (extent - 1) on every rank a non-singular kernel rank slides along
    const int pd = solid(n.args[0]);
    if (pd < 0) return;
    const PNode& pad = nodes[pd];
    if (!pad.func || pad.op != PAD || pad.dtype != FLOAT || pad.args.size() != 1) return;
    int64_t want_pad[rank_cap] = {0};
    ConvFuse cf;
    cf.dimg = true;
    for (int r = 0; r < rank_cap; ++r) cf.win[r] = 1;
    int prev = -1;
    for (int q = 0; q < rank_cap; ++q) {
      if (ks.at(q) == 1) continue;
      const int r = order[q];
      want_pad[r] = (int64_t)ks.at(q) - 1;
      if (q == 0) continue;
      if (r <= prev || r >= d) return;  // window coordinates must enumerate in the kernel's memory order
      prev = r;
      cf.win[r] = ks.at(q);
    }
    auto pads = eigen::unpack_dimpairs(*pad.func);
    for (int r = 0; r < rank_cap; ++r) {
      const int64_t lo = r < (int)pads.size() ? pads[r].first : 0, hi = r < (int)pads.size() ? pads[r].second : 0;
      if (lo != want_pad[r] || hi != want_pad[r]) return;
    }
    const PNode& sup = nodes[pad.args[0]];
    if (sup.dtype != FLOAT) return;
    const Shape ss = sup.shape;
    if ((int64_t)ss.at(d) != nout || (int64_t)n.shape.at(d) != 2 * p + 1) return;
    int64_t slice_n = 1;
    for (int r = 0; r < rank_cap; ++r) {
      if (r > d && (ss.at(r) != 1 || n.shape.at(r) != 1)) return;
      if (r >= d) { cf.img_shape[r] = 1; continue; }
      cf.img_shape[r] = (int64_t)ss.at(r) + cf.win[r] - 1;
      if ((int64_t)n.shape.at(r) != cf.img_shape[r]) return;
      cf.rows *= ss.at(r);
      cf.k *= cf.win[r];
      slice_n *= cf.img_shape[r];
    }
    cf.out_n = slice_n;
    cf.out_offset = (size_t)p * (size_t)slice_n * sizeof(float);
    // every reader — and every target, when the node is exposed through a view — must name exactly that slice
    if (n.exposed)
      for (iTensor* t : target_keys) {
        const PNode& tn = nodes[index.at(t)];
        if (tn.root == i && (tn.offset != cf.out_offset || tn.n != slice_n)) return;
      }
    for (size_t m = 0; m < nodes.size(); ++m) {
      const PNode& reader = nodes[m];
      if (!reader.func || reader.is_view) continue;
      for (int a : reader.args)
        if (nodes[a].root == i && (nodes[a].offset != cf.out_offset || nodes[a].n != slice_n)) return;
    }
    cf.pitch = (cf.k + 3) / 4 * 4;
    cf.img = rev0.args[0];  // the kernel itself
    cf.other = pad.args[0];
    cf.first = std::min({pd, ra, kr, i});
    tcr_gemm_desc& g = cf.gemm;
    std::memset(&g, 0, sizeof(g));
    g.m = cf.rows; g.n = cf.k; g.k = nout; g.batch = 1;
    g.a_sm = 1; g.a_sk = cf.rows;     // sup' [positions..., out]: out slowest
    g.b_sk = 1; g.b_sn = nout;        // kernel [out, (c,i,j)]: out fastest
    g.c_sm = cf.pitch; g.c_sn = 1;
    g.dtype = FLOAT;
    int through = -1;
    const int si = solid(pad.args[0]);
    if (si >= 0 && nodes[si].func && nodes[si].op == PERMUTE && nodes[si].args.size() == 1 && nodes[nodes[si].args[0]].dtype == FLOAT) {
      const RanksT sperm = full_order(*nodes[si].func);
      if (!sperm.empty()) {
        RanksT inv(rank_cap);
        for (int q = 0; q < rank_cap; ++q) inv[sperm[q]] = (RankT)q;
        if (d_first(ss, inv, d)) through = si;
      }
    }
    if (through >= 0) {
      cf.other = nodes[through].args[0];
      cf.first = std::min(cf.first, through);
      g.a_sm = nout; g.a_sk = 1;
    }
    if (assigned_between({cf.img, cf.other}, cf.first, i)) return;
    n.conv = (int)convs.size();
    convs.push_back(cf);
    ++skipped[pd];
    ++skipped[ra];
    chains.push_back({ra, kr});
    if (through >= 0) chains.push_back({pd, through});
  }

  void conv_step(Step& st, const ConvFuse& cf, int out_node, int bias_node, int epi, int act) {
    st.ew = false;
    st.conv_fused = true;
    st.conv_grad = cf.grad;
    st.out_node = out_node;
    st.gemm = cf.gemm;
    st.gemm.epilogue = epi;
    st.gemm.activation = act;
    for (int r = 0; r < rank_cap; ++r) { st.conv_img[r] = cf.img_shape[r]; st.conv_win[r] = cf.win[r]; }
    st.conv_pitch = cf.pitch;
    st.conv_rows = cf.rows;
    st.conv_dimg = cf.dimg;
    st.conv_out_offset = cf.out_offset;
    for (int a : {cf.dimg ? cf.other : cf.img, cf.dimg ? cf.img : cf.other, bias_node}) {
      if (a < 0) continue;
      st.in_nodes.push_back(nodes[a].root);
      st.in_offsets.push_back(nodes[a].offset);
    }
  }

  // dense layers: act(CONTRACT(x, W) + EXTEND(b)) becomes one GEMM launch with a bias (+ activation)
  // epilogue when the product and the sum have no other reader (cfg/tenncor/layer.yml:636-656,
  // nn.yml:14-47: the backward pass reads the activation's output, not the pre-activation)
  void fuse_gemm_epilogues() {
    if (std::getenv("TCR_NO_GEMM_EPILOGUE")) return;
    // weight gradients: PERMUTE{1,0}(CONTRACT(sup, x)) (backprop.hpp:269-359) = the product with
    // its operands swapped; both operand majors are native to the GEMM kernels
    for (size_t i = 0; i < nodes.size(); ++i) {
      PNode& n = nodes[i];
      if (!n.func || n.is_view || n.op != PERMUTE || n.args.size() != 1) continue;
      const int gi = nodes[n.args[0]].root;
      PNode& g = nodes[gi];
      if (nodes[n.args[0]].offset != 0 || !g.func || (g.op != CONTRACT && g.op != MATMUL) || g.exposed || g.inlined || g.consumers.size() != 1) continue;
      auto gop = dynamic_cast<DevOp*>(g.holder);
      if (!gop || !gop->gemm() || gop->gemm()->batch != 1) continue;
      const tcr_gemm_desc& d = *gop->gemm();
      if (d.c_sn != 1 || d.c_sm != d.n) continue;
      RanksT order = eigen::unpack_ranks(*n.func);
      if (order.size() < 2 || order[0] != 1 || order[1] != 0) continue;
      bool rest_identity = true;
      for (size_t r = 2; r < order.size() && r < rank_cap; ++r)
        if (order[r] != r) rest_identity = false;
      bool two_d = (int64_t)g.shape.at(0) == d.n && (int64_t)g.shape.at(1) == d.m;
      for (int r = 2; r < rank_cap; ++r)
        if (g.shape.at(r) != 1) two_d = false;
      if (!rest_identity || !two_d) continue;
      bool crosses = false;
      for (int arg : g.args) {
        auto it = assign_pos.find(nodes[arg].root);
        if (it == assign_pos.end()) continue;
        for (int pos : it->second)
          if (pos > gi && pos < (int)i) crosses = true;
      }
      if (crosses) continue;
      n.gemm_src = gi;
      n.gemm_transposed = true;
      g.inlined = true;
    }
    for (size_t i = 0; i < nodes.size(); ++i) {
      PNode& n = nodes[i];
      if (!n.func || n.is_view || n.op != ADD || n.args.size() != 2 || n.dtype != FLOAT) continue;
      for (int k = 0; k < 2; ++k) {
        const int gi = nodes[n.args[k]].root, oi = nodes[n.args[1 - k]].root;
        PNode& g = nodes[gi];
        PNode& o = nodes[oi];
        const bool from_conv = g.conv >= 0 && !convs[g.conv].grad && !convs[g.conv].dimg;  // conv2d composite: same epilogue on its GEMM
        if (!g.func || (!from_conv && g.op != CONTRACT && g.op != MATMUL) || g.exposed || g.inlined || g.consumers.size() != 1) continue;
        if (nodes[n.args[k]].offset != 0 || nodes[n.args[1 - k]].offset != 0 || !(g.shape == n.shape)) continue;
        auto gop = dynamic_cast<DevOp*>(g.holder);
        if (!from_conv && (!gop || !gop->gemm() || gop->gemm()->batch != 1)) continue;
        if (!o.is_extend || o.has_scalar || o.exposed || !(o.shape == n.shape) || o.consumers.size() != 1) continue;
        const PNode& b = nodes[o.args[0]];
        // which GEMM index does the bias follow? C is row-major [m][n]: n spans the leading ranks
        const tcr_gemm_desc& d = from_conv ? convs[g.conv].gemm : *gop->gemm();
        if (d.c_sn != 1 || d.c_sm != d.n) continue;
        int epi = 0;
        {
          int64_t lead = 1;
          int r = 0;
          for (; r < rank_cap && lead < d.n; ++r) lead *= n.shape.at(r);
          if (lead != d.n) continue;
          bool n_only = true, m_only = true;  // bias extents only inside / only outside the n ranks
          for (int q = 0; q < rank_cap; ++q) {
            const auto e = b.shape.at(q);
            if (q < r) { if (e != n.shape.at(q)) n_only = false; if (e != 1) m_only = false; }
            else { if (e != 1) n_only = false; if (e != n.shape.at(q)) m_only = false; }
          }
          if (n_only && b.n == d.n) epi = TCR_EPI_BIAS_N;
          else if (m_only && b.n == d.m) epi = TCR_EPI_BIAS_M;
          else continue;
        }
        // the GEMM moves to the position of the fused node: it must not cross an update of what it reads
        int final_node = (int)i, act = 0;
        if (!n.exposed && n.consumers.size() == 1) {
          PNode& c = nodes[n.consumers[0]];
          if ((c.op == SIGMOID || c.op == TANH) && c.args.size() == 1 && nodes[c.args[0]].root == (int)i && nodes[c.args[0]].offset == 0 &&
              c.dtype == FLOAT && !c.is_view) {
            final_node = n.consumers[0];
            act = c.op;
          }
        }
        bool crosses = false;
        if (from_conv) crosses = assigned_between({convs[g.conv].img, convs[g.conv].other}, convs[g.conv].first, final_node);
        else
        for (int arg : g.args) {
          auto it = assign_pos.find(nodes[arg].root);
          if (it == assign_pos.end()) continue;
          for (int pos : it->second)
            if (pos > gi && pos < final_node) crosses = true;
        }
        {
          auto it = assign_pos.find(nodes[o.args[0]].root);
          if (it != assign_pos.end())
            for (int pos : it->second)
              if (pos > oi && pos < final_node) crosses = true;
        }
        if (crosses) continue;
        PNode& f = nodes[final_node];
        f.gemm_src = gi;
        f.gemm_bias = o.args[0];
        f.gemm_epi = epi;
        f.gemm_act = act;
        f.is_ew = false;
        g.inlined = true;
        o.inlined = true;
        if (final_node != (int)i) n.inlined = true;
        break;
      }
    }
  }

  void fuse() {
    for (size_t i = 0; i < nodes.size(); ++i)
      if (nodes[i].func && is_assign(nodes[i].op)) assign_pos[nodes[nodes[i].args[0]].root].push_back((int)i);
    // consumers (on storage roots); EXTENDs are looked through later
    for (size_t i = 0; i < nodes.size(); ++i) {
      PNode& n = nodes[i];
      if (!n.func || n.is_view) continue;
      for (int a : n.args) nodes[nodes[a].root].consumers.push_back((int)i);
    }
    fuse_convs();
    fuse_gemm_epilogues();
    // non-EW consumers need real buffers: EXTEND operands get materialised. An EW node that cannot
    // be lowered on its own (too many operands / broadcast segments) is demoted to its holder's
    // kernel, which changes what its operands need — iterate to a fixed point.
    for (bool changed = true; changed;) {
      changed = false;
      for (auto& n : nodes) n.needs_mat = false;
      for (auto& cf : convs)
        for (int a : {cf.img, cf.other}) {
          PNode& r = nodes[nodes[a].root];
          if (r.is_extend) r.needs_mat = true;
        }
      for (size_t i = 0; i < nodes.size(); ++i) {
        PNode& n = nodes[i];
        if (!n.func || n.is_view) continue;
        if (n.gemm_src >= 0 && nodes[n.gemm_src].conv < 0) {
          // a fused GEMM step reads the product's operands (and its bias) as real buffers
          for (int a : nodes[n.gemm_src].args) {
            PNode& r = nodes[nodes[a].root];
            if (r.is_extend) r.needs_mat = true;
          }
          if (n.gemm_bias >= 0 && nodes[nodes[n.gemm_bias].root].is_extend) nodes[nodes[n.gemm_bias].root].needs_mat = true;
        }
        if (n.inlined || n.gemm_src >= 0) continue;  // absorbed into a GEMM epilogue: reads nothing itself
        bool ew_consumer = n.is_ew;
        for (size_t k = 0; k < n.args.size(); ++k) {
          PNode& a = nodes[nodes[n.args[k]].root];
          if (!a.is_extend) continue;
          bool as_bcast = ew_consumer && a.shape == n.shape && nodes[n.args[k]].offset == 0 && !(is_assign(n.op) && k == 0);
          if (!as_bcast && !a.has_scalar) a.needs_mat = true;
          if (!as_bcast && a.has_scalar && !ew_consumer) a.needs_mat = true;
        }
      }
      for (size_t i = 0; i < nodes.size(); ++i)
        if (nodes[i].is_extend && nodes[i].exposed) nodes[i].needs_mat = true;
      for (size_t i = 0; i < nodes.size(); ++i) {
        PNode& x = nodes[i];
        if (!x.is_ew || is_assign(x.op)) continue;
        x.region = (int)i;
        Step scratch;
        bool ok = try_region((int)i, scratch);
        x.region = -1;
        if (!ok) {
          x.is_ew = false;
          changed = true;
        }
      }
    }
    // greedy inlining in evaluation order
    for (size_t i = 0; i < nodes.size(); ++i) {
      PNode& x = nodes[i];
      if (!x.is_ew) continue;
      x.region = (int)i;
      members[(int)i] = {(int)i};
      Step scratch;
      for (size_t k = 0; k < x.args.size(); ++k) {
        if (is_assign(x.op) && k == 0) continue;  // the variable itself
        if (x.op == CAST) continue;
        int ai = x.args[k];
        if (nodes[ai].offset != 0) continue;
        int a = nodes[ai].root;
        PNode& an = nodes[a];
        if (!an.is_ew || an.inlined || an.exposed || is_assign(an.op) || an.region != a) continue;
        if (an.consumers.size() != 1 || an.n != x.n || an.dtype != x.dtype) continue;
        // a CAST joins its consumer's program when that program computes in the CAST's output type: its operand is never a
        // member (see above), so the conversion is exactly the typed load of that operand
        if (crosses_assign(a, (int)i)) continue;
        // tentatively merge
        std::vector<int> moved = members[a];
        for (int m : moved) nodes[m].region = (int)i;
        an.inlined = true;
        if (try_region((int)i, scratch)) {
          auto& mine = members[(int)i];
          mine.insert(mine.end(), moved.begin(), moved.end());
          members.erase(a);
        } else {
          for (int m : moved) nodes[m].region = a;
          an.inlined = false;
        }
      }
    }
  }

  // ---------------------------------------------------------------- steps + buffers
  void* alloc_owned(size_t bytes) {
    void* p = memory->allocate(bytes);
    owned.push_back({p, bytes});
    return p;
  }

  void build_steps() {
    for (size_t i = 0; i < nodes.size(); ++i) {
      PNode& n = nodes[i];
      if (!n.func || n.is_view || n.inlined) continue;
      if (n.is_extend && !n.needs_mat) continue;
      Step st;
      if (n.gemm_src >= 0 && nodes[n.gemm_src].conv >= 0) {
        conv_step(st, convs[nodes[n.gemm_src].conv], (int)i, n.gemm_bias, n.gemm_epi, n.gemm_act);
        n.step = (int)steps.size();
        steps.push_back(std::move(st));
        continue;
      }
      if (n.conv >= 0) {
        conv_step(st, convs[n.conv], (int)i, -1, TCR_EPI_NONE, 0);
        n.step = (int)steps.size();
        steps.push_back(std::move(st));
        continue;
      }
      if (n.gemm_src >= 0) {
        PNode& g = nodes[n.gemm_src];
        auto gop = dynamic_cast<DevOp*>(g.holder);
        st.ew = false;
        st.holder = gop;
        st.out_node = (int)i;
        st.gemm_fused = true;
        st.gemm = *gop->gemm();
        st.gemm.epilogue = n.gemm_epi;
        st.gemm.activation = n.gemm_act;
        if (n.gemm_transposed) {
          // C^T (n x m, row-major) = B^T (n x k) * A^T (k x m)
          const tcr_gemm_desc d = st.gemm;
          st.gemm.m = d.n; st.gemm.n = d.m;
          st.gemm.a_sm = d.b_sn; st.gemm.a_sk = d.b_sk;
          st.gemm.b_sk = d.a_sk; st.gemm.b_sn = d.a_sm;
          st.gemm.c_sm = d.m; st.gemm.c_sn = 1;
          for (int a : {g.args[1], g.args[0]}) {
            st.in_nodes.push_back(nodes[a].root);
            st.in_offsets.push_back(nodes[a].offset);
          }
        } else
        for (int a : {g.args[0], g.args[1], n.gemm_bias}) {
          st.in_nodes.push_back(nodes[a].root);
          st.in_offsets.push_back(nodes[a].offset);
        }
        n.step = (int)steps.size();
        steps.push_back(std::move(st));
        continue;
      }
      if (n.is_ew && try_region((int)i, st)) {
        n.step = (int)steps.size();
        steps.push_back(std::move(st));
        continue;
      }
      if (is_assign(n.op)) global::fatalf("planner: cannot lower %s", n.tens->to_string().c_str());
      auto op = dynamic_cast<DevOp*>(n.holder);
      if (!op) global::fatalf("planner: %s has no launchable holder", n.tens->to_string().c_str());
      st.ew = false;
      st.holder = op;
      st.out_node = (int)i;
      for (int a : n.args) {
        if (n.op == IDENTITY && !st.in_nodes.empty()) break;  // all-reduce marker: operational deps are not data
        st.in_nodes.push_back(nodes[a].root);
        st.in_offsets.push_back(nodes[a].offset);
      }
      n.step = (int)steps.size();
      steps.push_back(std::move(st));
    }
  }

  // ---------------------------------------------------------------- recurrent-step fusion (step level)
  // The reference unrolls layer.lstm / layer.gru / layer.rnn over time (cfg/tenncor/layer.yml:678-813) and teq::derive sums
  // the per-step contributions of every shared weight with one n-ary ADD (internal/teq/src/derive.cpp:49-51). After
  // build_steps() that is, per time step, four small GEMMs on the same operand, a CONCAT, four more GEMMs + an ADD + a SLICE for
  // the gradient reaching h_{t-1}, and four weight-gradient GEMMs + four bias reductions whose results meet in 128-operand ADDs.
  // This pass rewrites the STEP list (values are unchanged; the oracle still evaluates the unfused graph):
  //   (1) sibling products  act(A . W_g + b_g), same A          -> ONE grouped launch (tcr_gemm_grouped); A = CONCAT(x, h)
  //                                                                becomes two K-segments and the CONCAT step disappears when
  //                                                                nobody else reads it
  //   (2) ADD of products   sum_g (dpre_g . W_g^T) [+ SLICE]      -> ONE K-segmented launch producing only the sliced columns
  //   (3) n-ary ADD of T weight-gradient products A_t^T . B_t     -> the operands are PLACED back to back (stacks) and the sum
  //                                                                is ONE product with K = T . k
  //   (4) n-ary ADD of T bias reductions                          -> ONE reduction over the stack
  //   (5) n-ary CONCAT of T step outputs along the last rank      -> the producers write their slab of the result in place
  // Steps keep (storage root, version) read / write sets taken from the original order; the rewritten list is re-sorted by a
  // stable topological sort on those, so a merged step lands after everything any of its members waited for.
  struct Stack {
    std::vector<int> members;
    size_t slab = 0;
    void* base = nullptr;
  };
  std::vector<Stack> stacks;

  struct Acc {
    std::vector<std::pair<int, int>> rd, wr;
  };

  void step_reads(const Step& st, std::vector<int>& out) const {
    out.clear();
    if (st.kind == Step::BUCKET_FLUSH) return;
    if (st.ew) for (auto& in : st.inputs) out.push_back(in.node);
    else for (int in : st.in_nodes) out.push_back(in);
    for (int d : st.dep_nodes) out.push_back(d);
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
  }

  void step_writes(const Step& st, std::vector<int>& out) const {
    out.clear();
    if (st.kind == Step::BUCKET_FLUSH) return;
    out.push_back(st.out_node);
    for (int e : st.extra_outs) out.push_back(e);
    const PNode& o = nodes[st.out_node];
    if (is_assign(o.op)) out.push_back(nodes[o.args[0]].root);
  }

  struct GemmView {
    bool ok = false;
    tcr_gemm_desc d;
    int a = -1, b = -1, bias = -1;
    size_t a_off = 0, b_off = 0, bias_off = 0;
  };

  GemmView gemm_view(const Step& st) const {
    GemmView v;
    if (st.dead || st.kind != Step::NORMAL || st.ew || st.conv_fused || st.group || st.stack_reduce) return v;
    const PNode& o = nodes[st.out_node];
    if (st.gemm_fused) {
      v.d = st.gemm;
      if (st.in_nodes.size() > 2) { v.bias = st.in_nodes[2]; v.bias_off = st.in_offsets[2]; }
    } else {
      if (!st.holder || !st.holder->gemm() || (o.op != CONTRACT && o.op != MATMUL) || st.in_nodes.size() != 2) return v;
      v.d = *st.holder->gemm();
      v.d.epilogue = TCR_EPI_NONE;
      v.d.activation = 0;
    }
    if (v.d.dtype != FLOAT || v.d.batch != 1 || v.d.accumulate) return v;
    v.a = st.in_nodes[0]; v.a_off = st.in_offsets[0];
    v.b = st.in_nodes[1]; v.b_off = st.in_offsets[1];
    v.ok = true;
    return v;
  }

  // operands 0..T-1 laid out back to back: all views of one buffer at base + i * slab, or distinct plan-owned results that can be
  // (or already are) placed consecutively in a stack. check-only unless `commit`.
  bool stack_operands(const std::vector<std::pair<int, size_t>>& refs, size_t slab, bool commit, int& base_node, size_t& base_off) {
    if (refs.empty()) return false;
    bool same_root = true;
    for (size_t i = 0; i < refs.size(); ++i)
      if (refs[i].first != refs[0].first || refs[i].second != refs[0].second + i * slab) same_root = false;
    if (same_root) { base_node = refs[0].first; base_off = refs[0].second; return true; }
    std::set<int> seen;
    int sid = -2;
    for (size_t i = 0; i < refs.size(); ++i) {
      const PNode& n = nodes[refs[i].first];
      if (refs[i].second != 0 || !n.func || n.exposed || n.is_view || is_assign(n.op) || n.step < 0 || n.bucket_slot >= 0 || n.conv >= 0 ||
          n.alias_stack >= 0 || (size_t)n.n * type_size(n.dtype) != slab || !seen.insert(refs[i].first).second)
        return false;
      if (i == 0) sid = n.stack;
      if (n.stack != sid) return false;
      if (sid >= 0 && n.stack_pos != nodes[refs[0].first].stack_pos + (int)i) return false;
    }
    if (sid >= 0) {
      base_node = stacks[sid].members[0];
      base_off = (size_t)nodes[refs[0].first].stack_pos * slab;
      return true;
    }
    if (commit) {
      Stack st;
      st.slab = slab;
      for (size_t i = 0; i < refs.size(); ++i) {
        st.members.push_back(refs[i].first);
        nodes[refs[i].first].stack = (int)stacks.size();
        nodes[refs[i].first].stack_pos = (int)i;
      }
      stacks.push_back(std::move(st));
    }
    base_node = refs[0].first;
    base_off = 0;
    return true;
  }

  void fuse_recurrent_steps() {
    if (std::getenv("TCR_NO_RNN_FUSE") || precision == TCR_GEMM_EXACT) return;
    const int ns = (int)steps.size();
    if (ns < 3) return;
    const std::vector<Step> original = steps;
    std::vector<int> original_step_of(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) original_step_of[i] = nodes[i].step;
    auto restore = [&] {
      steps = original;
      for (size_t i = 0; i < nodes.size(); ++i) { nodes[i].step = original_step_of[i]; nodes[i].stack = -1; nodes[i].alias_stack = -1; }
      stacks.clear();
    };
    // ---- read / write sets with versions, from the original order
    std::vector<Acc> acc(ns);
    std::vector<std::vector<int>> readers(nodes.size());
    {
      std::unordered_map<int, int> ver;
      std::vector<int> rd, wr;
      for (int s = 0; s < ns; ++s) {
        steps[s].pos = s;
        step_reads(steps[s], rd);
        step_writes(steps[s], wr);
        for (int r : rd) { acc[s].rd.push_back({r, ver[r]}); readers[r].push_back(s); }
        for (int w : wr) acc[s].wr.push_back({w, ++ver[w]});
      }
    }
    auto merge_acc = [&](int into, int from) {
      acc[into].rd.insert(acc[into].rd.end(), acc[from].rd.begin(), acc[from].rd.end());
      acc[into].wr.insert(acc[into].wr.end(), acc[from].wr.begin(), acc[from].wr.end());
    };
    auto drop_node = [&](int s, int node) {  // `node` became internal to step s (or virtual)
      auto& a = acc[s];
      a.rd.erase(std::remove_if(a.rd.begin(), a.rd.end(), [&](const std::pair<int, int>& x) { return x.first == node; }), a.rd.end());
      a.wr.erase(std::remove_if(a.wr.begin(), a.wr.end(), [&](const std::pair<int, int>& x) { return x.first == node; }), a.wr.end());
    };
    auto live_readers = [&](int node) {
      std::vector<int> out;
      for (int r : readers[node])
        if (!steps[r].dead) out.push_back(r);
      return out;
    };
    auto producer = [&](int node) -> int {
      const PNode& n = nodes[node];
      if (!n.func || n.step < 0 || n.step >= ns || steps[n.step].dead || steps[n.step].out_node != node) return -1;
      return n.step;
    };
    auto plain_f32 = [&](int node) { return nodes[node].dtype == FLOAT; };
    int n_fused = 0;

    // ================= (0) activation-derivative factor folded into the product that feeds it =================
    // backprop.hpp:136-142: d SIGMOID = MUL(MUL(s, SUB(1, s)), sup), d TANH = MUL(SUB(1, SQUARE(t)), sup). When `sup` is a small-K
    // product (the input gradient of a narrow dense layer: K = its output width) the factor is applied while the product is
    // written: one pass over the [batch x hidden] gradient instead of write, read twice, write.
    for (int s = 0; s < ns; ++s) {
      Step& e = steps[s];
      if (e.dead || !e.ew || e.kind != Step::NORMAL || e.prog.n_outputs != 1 || e.prog.n_inputs != 2) continue;
      const int n0 = e.out_node;
      const PNode& o = nodes[n0];
      auto mem = members.find(n0);
      if (o.op != MUL || o.args.size() != 2 || o.dtype != FLOAT || is_assign(o.op) || mem == members.end() || mem->second.size() != 3) continue;
      const PNode& loc = nodes[nodes[o.args[0]].root];
      const PNode& sup = nodes[o.args[1]];
      if (!loc.func || !loc.inlined || loc.region != n0 || sup.offset != 0 || nodes[sup.root].n != o.n) continue;
      int post = 0, aux = -1;
      auto is_one = [&](int a) { return nodes[a].has_scalar && nodes[a].scalar == 1.0; };
      auto plain = [&](int a) { const PNode& x = nodes[a]; return (!x.has_scalar && x.offset == 0 && x.n == o.n && x.dtype == FLOAT && !x.is_extend) ? x.root : -1; };
      if (loc.op == MUL && loc.args.size() == 2) {  // s * (1 - s)
        const int sv = plain(loc.args[0]);
        const PNode& sb = nodes[nodes[loc.args[1]].root];
        if (sv >= 0 && sb.func && sb.op == SUB && sb.inlined && sb.region == n0 && sb.args.size() == 2 && is_one(sb.args[0]) && plain(sb.args[1]) == sv) { post = TCR_POST_MUL_DSIGMOID; aux = sv; }
      } else if (loc.op == SUB && loc.args.size() == 2 && is_one(loc.args[0])) {  // 1 - t^2
        const PNode& sq = nodes[nodes[loc.args[1]].root];
        if (sq.func && sq.op == SQUARE && sq.inlined && sq.region == n0 && sq.args.size() == 1) {
          const int tv = plain(sq.args[0]);
          if (tv >= 0) { post = TCR_POST_MUL_DTANH; aux = tv; }
        }
      }
      if (!post) continue;
      const int ps = producer(sup.root);
      if (ps < 0 || nodes[sup.root].exposed || live_readers(sup.root).size() != 1) continue;
      GemmView v = gemm_view(steps[ps]);
      if (!v.ok || v.d.k > 16 || v.d.c_sn != 1 || v.d.c_sm != v.d.n || (v.d.n % 4) || v.d.n < 64 || v.d.m * v.d.n != o.n || v.d.post_op) continue;
      if (v.d.m * v.d.n < 4096) continue;
      Step g = steps[ps];
      g.gemm_fused = true;
      g.holder = nullptr;
      g.gemm = v.d;
      g.gemm.post_op = post;
      g.gemm_aux_slot = (int)g.in_nodes.size();
      g.in_nodes.push_back(aux);
      g.in_offsets.push_back(0);
      g.out_node = n0;
      g.pos = e.pos;
      merge_acc(s, ps);
      drop_node(s, sup.root);
      steps[ps].dead = true;
      steps[s] = std::move(g);
      ++n_fused;
    }

    // ================= (0b) loss chains: REDUCE_SUM over every rank of an elementwise result [/ constant] =================
    // cfg/tenncor/loss.yml:21-39 -> core.yml:1090-1096: mean_squared = DIV(REDUCE_SUM(SQUARE(SUB(a, b))), count). The elementwise
    // program sums its own output (deterministic block-order partials) and applies the scalar division: one launch, and the
    // [batch x out] intermediate is never written.
    if (!std::getenv("TCR_NO_LOSS_FUSE"))
    for (int s = 0; s < ns; ++s) {
      Step& r = steps[s];
      if (r.dead || r.ew || r.kind != Step::NORMAL || r.gemm_fused || r.conv_fused || r.group || r.stack_reduce || !r.holder) continue;
      const int rnode = r.out_node;
      const PNode& rn = nodes[rnode];
      if (rn.op != REDUCE_SUM || rn.n != 1 || (rn.dtype != FLOAT && rn.dtype != DOUBLE) || r.in_nodes.size() != 1 || r.in_offsets[0] != 0) continue;
      const int x = r.in_nodes[0];
      const int ps = producer(x);
      if (ps < 0 || nodes[x].exposed || nodes[x].stack >= 0 || live_readers(x).size() != 1 || nodes[rn.args[0]].n != nodes[x].n) continue;
      const Step& e = steps[ps];
      if (!e.ew || e.kind != Step::NORMAL || e.ew_reduce || e.prog.n_outputs != 1 || is_assign(nodes[x].op) || nodes[x].dtype != rn.dtype || e.prog.dtype != (int)rn.dtype ||
          e.prog.outputs[0].dtype != (int)rn.dtype)
        continue;
      Step f = e;
      f.ew_reduce = true;
      f.out_node = rnode;
      f.pos = r.pos;
      merge_acc(s, ps);
      drop_node(s, x);
      steps[ps].dead = true;
      steps[s] = std::move(f);
      ++n_fused;
      // the scalar that follows: DIV / MUL of the sum by a constant (reduce_mean's count)
      if (rn.exposed) continue;
      const std::vector<int> rd = live_readers(rnode);
      if (rd.size() != 1) continue;
      const int sd = rd[0];
      const Step& d = steps[sd];
      if (sd <= s || !d.ew || d.kind != Step::NORMAL || d.ew_reduce || d.prog.n_outputs != 1 || d.prog.n_inputs != 1 || d.prog.n_instrs != 2 || d.inputs.size() != 1 ||
          d.inputs[0].node != rnode || d.inputs[0].offset != 0 || nodes[d.out_node].n != 1 || nodes[d.out_node].dtype != rn.dtype || is_assign(nodes[d.out_node].op) ||
          d.prog.dtype != (int)rn.dtype)
        continue;
      const tcr_ew_instr &i0 = d.prog.instrs[0], &i1 = d.prog.instrs[1];
      if (i0.op != TCR_EW_CONST || (i1.op != DIV && i1.op != MUL) || i1.a != 0 || i1.b != i0.dst || i0.dst == 0 || d.prog.outputs[0].reg != i1.dst) continue;
      Step g = steps[s];
      g.red_post = i1.op == DIV ? 1 : 2;
      g.red_imm = i0.imm;
      g.out_node = d.out_node;
      g.pos = d.pos;
      merge_acc(sd, s);
      drop_node(sd, rnode);
      steps[s].dead = true;
      steps[sd] = std::move(g);
    }

    // ================= (1) sibling products sharing their A operand -> grouped launch =================
    {
      struct Key {
        int a; size_t off; int64_t m, n, k, a_sm, b_sk;
        int b_version;  // products reading their weights before / after an in-place update belong to different launches
        bool operator<(const Key& o) const { return std::tie(a, off, m, n, k, a_sm, b_sk, b_version) < std::tie(o.a, o.off, o.m, o.n, o.k, o.a_sm, o.b_sk, o.b_version); }
      };
      std::map<Key, std::vector<int>> buckets;
      for (int s = 0; s < ns; ++s) {
        GemmView v = gemm_view(steps[s]);
        if (!v.ok) continue;
        const tcr_gemm_desc& d = v.d;
        if (d.m > 128 || d.n < 32 || d.k < 32) continue;                                  // the small-batch kernel
        if (d.a_sk != 1 || d.b_sn != 1 || d.c_sn != 1 || d.c_sm != d.n) continue;         // A row-major, B [k x n], C row-major
        if (d.epilogue != TCR_EPI_NONE && d.epilogue != TCR_EPI_BIAS_N) continue;
        if ((d.a_sm % 4) || (d.b_sk % 4) || (v.a_off % 16) || (v.b_off % 16)) continue;
        int b_version = 0;
        for (auto& r : acc[s].rd)
          if (r.first == v.b) b_version = r.second;
        buckets[Key{v.a, v.a_off, d.m, d.n, d.k, d.a_sm, d.b_sk, b_version}].push_back(s);
      }
      for (auto& kv : buckets) {
        std::vector<int>& all = kv.second;
        // the A operand as K-segments when it is a binary CONCAT along the fast rank that only products read
        const Key& key = kv.first;
        std::vector<std::pair<int, size_t>> segs = {{key.a, key.off}};
        std::vector<int64_t> seg_k = {key.k}, seg_pitch = {key.a_sm};
        int concat_step = -1;
        {
          const int ps = key.off == 0 ? producer(key.a) : -1;
          if (ps >= 0) {
            const Step& cs = steps[ps];
            const PNode& cn = nodes[key.a];
            if (!cs.ew && !cs.gemm_fused && cs.holder && cn.op == CONCAT && cs.in_nodes.size() == 2 && !cn.exposed && plain_f32(key.a) &&
                eigen::unpack_rank(*cn.func) == 0) {
              const PNode& x0 = nodes[cn.args[0]];
              const PNode& x1 = nodes[cn.args[1]];
              const int64_t k0 = x0.shape.at(0), k1 = x1.shape.at(0);
              bool ok = k0 + k1 == key.k && key.a_sm == key.k && x0.n == k0 * key.m && x1.n == k1 * key.m && (k0 % 4) == 0 && (k1 % 4) == 0 &&
                        (cs.in_offsets[0] % 16) == 0 && (cs.in_offsets[1] % 16) == 0 && x0.dtype == FLOAT && x1.dtype == FLOAT;
              if (ok) {
                segs = {{cs.in_nodes[0], cs.in_offsets[0]}, {cs.in_nodes[1], cs.in_offsets[1]}};
                seg_k = {k0, k1};
                seg_pitch = {k0, k1};
                concat_step = ps;
              }
            }
          }
        }
        if (all.size() < 2 && concat_step < 0) continue;
        for (size_t at = 0; at < all.size();) {
          const size_t left = all.size() - at;
          const int G = left >= 4 ? 4 : left >= 2 ? 2 : 1;
          if (G == 1 && concat_step < 0) { ++at; continue; }
          std::vector<int> mem(all.begin() + at, all.begin() + at + G);
          at += G;
          const int keep = mem.back();
          Step g;
          g.group = true;
          g.pos = steps[keep].pos;
          g.out_node = steps[keep].out_node;
          std::memset(&g.gd, 0, sizeof(g.gd));
          tcr_gemm_group_desc& gd = g.gd;
          gd.m = key.m; gd.n = key.n; gd.groups = G; gd.segments = (int)segs.size();
          gd.b_pitch = key.b_sk; gd.b_trans = 0; gd.out_pitch = key.n;
          for (size_t q = 0; q < segs.size(); ++q) {
            gd.seg_k[q] = seg_k[q];
            gd.a_pitch[q] = seg_pitch[q];
            g.in_nodes.push_back(segs[q].first);
            g.in_offsets.push_back(segs[q].second);
          }
          std::vector<GemmView> views;
          for (int m : mem) views.push_back(gemm_view(steps[m]));
          for (int gi = 0; gi < G; ++gi) {
            size_t koff = 0;
            for (size_t q = 0; q < segs.size(); ++q) {
              g.in_nodes.push_back(views[gi].b);
              g.in_offsets.push_back(views[gi].b_off + koff * (size_t)key.b_sk * sizeof(float));
              koff += (size_t)seg_k[q];
            }
          }
          for (int gi = 0; gi < G; ++gi) {
            gd.act[gi] = views[gi].d.activation;
            g.g_out_nodes[gi] = steps[mem[gi]].out_node;
            if (views[gi].d.epilogue == TCR_EPI_BIAS_N && views[gi].bias >= 0) {
              g.g_bias_slot[gi] = (int)g.in_nodes.size();
              g.in_nodes.push_back(views[gi].bias);
              g.in_offsets.push_back(views[gi].bias_off);
            }
            if (mem[gi] != keep) g.extra_outs.push_back(steps[mem[gi]].out_node);
          }
          gd.precision = TCR_GEMM_3XTF32;
          if (tcr_gemm_grouped_check(&gd) != TCR_OK) continue;
          // commit
          for (int m : mem)
            if (m != keep) { merge_acc(keep, m); steps[m].dead = true; }
          if (concat_step >= 0) {
            drop_node(keep, key.a);
            acc[keep].rd.insert(acc[keep].rd.end(), acc[concat_step].rd.begin(), acc[concat_step].rd.end());
          }
          g.dead = false;
          steps[keep] = std::move(g);
          ++n_fused;
        }
        if (concat_step >= 0) {  // nobody reads the concatenation any more: it is never built
          bool any = false;
          std::vector<int> rd;
          for (int r : live_readers(key.a)) {
            step_reads(steps[r], rd);
            if (std::find(rd.begin(), rd.end(), key.a) != rd.end()) any = true;
          }
          if (!any) steps[concat_step].dead = true;
        }
      }
    }

    // ================= (1b) LSTM cell update in the epilogue of the grouped gate launch =================
    // state = ADD(MUL(gate, in), MUL(state, forget)); hidden = MUL(state, output) (cfg/tenncor/layer.yml:758-760): when the four
    // operands are the four results of one grouped launch, the CTA that holds all gates of a unit computes both on the spot.
    for (int s = 0; s < ns; ++s) {
      if (steps[s].dead || !steps[s].group || steps[s].gd.groups != 4 || steps[s].gd.b_trans || steps[s].gd.cell) continue;
      int gate_of[4];
      for (int g = 0; g < 4; ++g) gate_of[g] = steps[s].g_out_nodes[g];
      auto which_gate = [&](int arg) {
        const PNode& a = nodes[arg];
        if (a.offset != 0) return -1;
        for (int g = 0; g < 4; ++g)
          if (a.root == gate_of[g]) return g;
        return -1;
      };
      // E1: the state update, an elementwise step reading gates of this launch
      int e1 = -1, e2 = -1, role_f = -1, role_o = -1, cprev = -2;
      size_t cprev_off = 0;
      int pair[2] = {-1, -1};
      for (int r : live_readers(gate_of[0])) {
        const Step& e = steps[r];
        if (!e.ew || e.kind != Step::NORMAL) continue;
        const int c = e.out_node;
        const PNode& cn = nodes[c];
        auto mem = members.find(c);
        if (cn.op != ADD || cn.args.size() != 2 || cn.dtype != FLOAT || mem == members.end() || mem->second.size() != 3) continue;
        const PNode &p = nodes[nodes[cn.args[0]].root], &q = nodes[nodes[cn.args[1]].root];
        if (!p.func || !q.func || p.op != MUL || q.op != MUL || p.args.size() != 2 || q.args.size() != 2 || !p.inlined || !q.inlined || p.region != c || q.region != c) continue;
        const int leaf[4] = {p.args[0], p.args[1], q.args[0], q.args[1]};
        int g4[4], n_gate = 0, other = -1;
        for (int k = 0; k < 4; ++k) {
          g4[k] = which_gate(leaf[k]);
          if (g4[k] >= 0) ++n_gate;
          else other = k;
        }
        if (n_gate != 3) continue;
        bool distinct = true;
        for (int a = 0; a < 4; ++a)
          for (int b = a + 1; b < 4; ++b)
            if (g4[a] >= 0 && g4[a] == g4[b]) distinct = false;
        if (!distinct) continue;
        const PNode& on = nodes[leaf[other]];
        if (on.has_scalar) {
          if (on.scalar != 0.0) continue;
          cprev = -1;  // zero state
        } else {
          if (on.n != cn.n || on.dtype != FLOAT || on.is_extend) continue;
          cprev = on.root;
          cprev_off = on.offset;
        }
        role_f = g4[other ^ 1];                       // the gate multiplied with the previous state
        pair[0] = g4[(other < 2) ? 2 : 0];            // the other product: candidate x input gate (commutative)
        pair[1] = g4[(other < 2) ? 3 : 1];
        role_o = 6 - role_f - pair[0] - pair[1];
        e1 = r;
        break;
      }
      if (e1 < 0) continue;
      const int c = steps[e1].out_node;
      for (int r : live_readers(c)) {
        const Step& e = steps[r];
        if (!e.ew || e.kind != Step::NORMAL) continue;
        const int h = e.out_node;
        const PNode& hn = nodes[h];
        auto mem = members.find(h);
        if (hn.op != MUL || hn.args.size() != 2 || hn.dtype != FLOAT || mem == members.end() || mem->second.size() != 1) continue;
        const PNode &x = nodes[hn.args[0]], &y = nodes[hn.args[1]];
        const bool cx = x.root == c && x.offset == 0, cy = y.root == c && y.offset == 0;
        if (cx == cy) continue;
        if (which_gate(cx ? hn.args[1] : hn.args[0]) != role_o) continue;
        e2 = r;
        break;
      }
      if (e2 < 0 || nodes[c].n != steps[s].gd.m * steps[s].gd.n) continue;
      const int h = steps[e2].out_node;
      Step& g = steps[s];
      tcr_gemm_group_desc trial = g.gd;
      trial.cell = 1;
      trial.role_cand = pair[0]; trial.role_in = pair[1]; trial.role_forget = role_f; trial.role_out = role_o;
      trial.state_pitch = trial.n;
      if (tcr_gemm_grouped_check(&trial) != TCR_OK) continue;
      g.gd = trial;
      if (cprev >= 0) {
        g.g_cprev_slot = (int)g.in_nodes.size();
        g.in_nodes.push_back(cprev);
        g.in_offsets.push_back(cprev_off);
      }
      g.g_c_node = c;
      g.g_h_node = h;
      merge_acc(s, e1);
      merge_acc(s, e2);
      steps[e1].dead = true;
      steps[e2].dead = true;
      // a gate nobody else reads is not stored
      g.extra_outs.clear();
      g.out_node = h;
      g.extra_outs.push_back(c);
      for (int k = 0; k < 4; ++k) {
        bool read = nodes[gate_of[k]].exposed;
        std::vector<int> rd;
        for (int r : live_readers(gate_of[k])) {
          if (r == s) continue;
          step_reads(steps[r], rd);
          if (std::find(rd.begin(), rd.end(), gate_of[k]) != rd.end()) read = true;
        }
        if (read) g.extra_outs.push_back(gate_of[k]);
        else { g.g_out_nodes[k] = -1; drop_node(s, gate_of[k]); }
      }
      g.pos = std::max(g.pos, steps[e2].pos);
      ++n_fused;
    }

    // ================= (2) ADD of products (+ SLICE of the sum) -> one K-segmented launch =================
    for (int s = 0; s < ns; ++s) {
      Step& e = steps[s];
      if (e.dead || !e.ew || e.kind != Step::NORMAL || e.prog.n_outputs != 1) continue;
      const int nin = e.prog.n_inputs;
      if (nin < 2 || nin > 4 || e.prog.n_instrs != nin - 1 || nodes[e.out_node].dtype != FLOAT || is_assign(nodes[e.out_node].op)) continue;
      bool pure = true;
      for (int k = 0; k < e.prog.n_instrs; ++k)
        if (e.prog.instrs[k].op != TCR_EW_ADD) pure = false;
      std::vector<int> prods;
      std::vector<GemmView> views;
      for (int k = 0; k < nin && pure; ++k) {
        const InputRef& in = e.inputs[k];
        const int ps = (in.mask == 0 && in.offset == 0 && in.dtype == FLOAT) ? producer(in.node) : -1;
        if (ps < 0 || nodes[in.node].exposed || live_readers(in.node).size() != 1 || std::find(prods.begin(), prods.end(), ps) != prods.end()) { pure = false; break; }
        GemmView v = gemm_view(steps[ps]);
        if (!v.ok || v.d.epilogue != TCR_EPI_NONE || v.d.activation) { pure = false; break; }
        const tcr_gemm_desc& d = v.d;
        if (d.m > 128 || d.n < 32 || d.a_sk != 1 || d.b_sk != 1 || d.c_sn != 1 || d.c_sm != d.n || (d.a_sm % 4) || (d.b_sn % 4) || (v.a_off % 16) || (v.b_off % 16)) { pure = false; break; }
        if (!views.empty() && (d.m != views[0].d.m || d.n != views[0].d.n || d.b_sn != views[0].d.b_sn)) { pure = false; break; }
        prods.push_back(ps);
        views.push_back(v);
      }
      if (!pure || (int)views.size() != nin) continue;
      // every input must feed the sum exactly once: n - 1 ADDs over n distinct inputs is a sum of all of them iff each register 0..n-1 is consumed
      int keep = s;
      int64_t n_cols = views[0].d.n, col0 = 0;
      int out_node = e.out_node;
      // SLICE of the fast rank as the only reader: compute only those columns
      {
        std::vector<int> rs = live_readers(e.out_node);
        if (rs.size() == 1 && !nodes[e.out_node].exposed) {
          const Step& sl = steps[rs[0]];
          const PNode& sn = nodes[sl.out_node];
          if (!sl.ew && !sl.gemm_fused && !sl.group && sl.holder && sn.op == SLICE && sl.in_nodes.size() == 1 && sl.in_offsets[0] == 0 && sn.dtype == FLOAT) {
            auto cuts = eigen::unpack_dimpairs(*sn.func);
            const Shape whole = nodes[e.out_node].shape;
            bool only_fast = true;
            for (size_t r = 1; r < rank_cap; ++r) {
              if (r < cuts.size() && (cuts[r].first != 0 || std::min<int64_t>(cuts[r].second, whole.at(r)) != (int64_t)whole.at(r))) only_fast = false;
            }
            if (only_fast && !cuts.empty() && (int64_t)whole.at(0) == n_cols) {
              const int64_t lo = std::min<int64_t>(cuts[0].first, whole.at(0) - 1);
              const int64_t len = std::min<int64_t>(cuts[0].second, whole.at(0) - lo);
              if ((int64_t)sn.shape.at(0) == len && len >= 32 && (len % 4) == 0 && ((lo * views[0].d.b_sn * 4) % 16) == 0) {
                col0 = lo; n_cols = len; keep = rs[0]; out_node = sl.out_node;
              }
            }
          }
        }
      }
      Step g;
      g.group = true;
      g.pos = steps[keep].pos;
      g.out_node = out_node;
      std::memset(&g.gd, 0, sizeof(g.gd));
      tcr_gemm_group_desc& gd = g.gd;
      gd.m = views[0].d.m; gd.n = n_cols; gd.groups = 1; gd.segments = nin;
      gd.b_pitch = views[0].d.b_sn; gd.b_trans = 1; gd.out_pitch = n_cols;
      for (int k = 0; k < nin; ++k) {
        gd.seg_k[k] = views[k].d.k;
        gd.a_pitch[k] = views[k].d.a_sm;
        g.in_nodes.push_back(views[k].a);
        g.in_offsets.push_back(views[k].a_off);
      }
      for (int k = 0; k < nin; ++k) {
        g.in_nodes.push_back(views[k].b);
        g.in_offsets.push_back(views[k].b_off + (size_t)col0 * (size_t)gd.b_pitch * sizeof(float));
      }
      g.g_out_nodes[0] = out_node;
      gd.precision = TCR_GEMM_3XTF32;
      if (tcr_gemm_grouped_check(&gd) != TCR_OK) continue;
      const int sum_node = e.out_node;
      if (keep != s) { merge_acc(keep, s); steps[s].dead = true; }
      for (int ps : prods) { merge_acc(keep, ps); steps[ps].dead = true; }
      for (int ps : prods) drop_node(keep, steps[ps].out_node);
      if (keep != s) drop_node(keep, sum_node);
      steps[keep] = std::move(g);
      ++n_fused;
    }

    // ================= (3) n-ary ADD of weight-gradient products -> operands stacked, ONE product =================
    // ================= (4) n-ary ADD of reductions               -> ONE reduction over the stack    =================
    for (int pass = 0; pass < 2; ++pass)
    for (int s = 0; s < ns; ++s) {
      Step& x = steps[s];
      if (x.dead || x.ew || x.gemm_fused || x.group || x.stack_reduce || x.kind != Step::NORMAL || !x.holder) continue;
      const PNode& xn = nodes[x.out_node];
      if (xn.op != ADD || xn.dtype != FLOAT || x.in_nodes.size() < 3) continue;
      const int T = (int)x.in_nodes.size();
      std::vector<int> prods;
      bool ok = true;
      for (int i = 0; i < T && ok; ++i) {
        const int ps = x.in_offsets[i] == 0 ? producer(x.in_nodes[i]) : -1;
        if (ps < 0 || nodes[x.in_nodes[i]].exposed || live_readers(x.in_nodes[i]).size() != 1 || std::find(prods.begin(), prods.end(), ps) != prods.end()) ok = false;
        else prods.push_back(ps);
      }
      if (!ok) continue;
      if (pass == 0) {
        std::vector<GemmView> views;
        for (int ps : prods) {
          GemmView v = gemm_view(steps[ps]);
          if (!v.ok || v.d.epilogue != TCR_EPI_NONE || v.d.activation) { ok = false; break; }
          if (!views.empty()) {
            const tcr_gemm_desc &d = v.d, &f = views[0].d;
            if (d.m != f.m || d.n != f.n || d.k != f.k || d.a_sm != f.a_sm || d.a_sk != f.a_sk || d.b_sk != f.b_sk || d.b_sn != f.b_sn || d.c_sm != f.c_sm || d.c_sn != f.c_sn) { ok = false; break; }
          }
          views.push_back(v);
        }
        if (!ok) continue;
        // canonical order: by the A operand's position in the graph (forward time order)
        std::vector<int> order(T);
        for (int i = 0; i < T; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int p, int q) { return std::make_pair(views[p].a, views[p].a_off) < std::make_pair(views[q].a, views[q].a_off); });
        const tcr_gemm_desc& f = views[0].d;
        const size_t a_slab = (size_t)f.k * (size_t)f.a_sk * sizeof(float), b_slab = (size_t)f.k * (size_t)f.b_sk * sizeof(float);
        if (f.a_sk < f.m || f.b_sk < f.n) continue;  // K must be the slow extent of both operands
        std::vector<std::pair<int, size_t>> ar, br;
        for (int i : order) { ar.push_back({views[i].a, views[i].a_off}); br.push_back({views[i].b, views[i].b_off}); }
        int an, bn;
        size_t ao, bo;
        if (!stack_operands(ar, a_slab, false, an, ao) || !stack_operands(br, b_slab, false, bn, bo)) continue;
        stack_operands(ar, a_slab, true, an, ao);
        stack_operands(br, b_slab, true, bn, bo);
        Step g;
        g.gemm_fused = true;
        g.pos = x.pos;
        g.out_node = x.out_node;
        g.gemm = f;
        g.gemm.k = f.k * T;
        g.in_nodes = {an, bn};
        g.in_offsets = {ao, bo};
        for (auto& r : ar) g.dep_nodes.push_back(r.first);
        for (auto& r : br) g.dep_nodes.push_back(r.first);
        for (int ps : prods) { merge_acc(s, ps); drop_node(s, steps[ps].out_node); steps[ps].dead = true; }
        steps[s] = std::move(g);
        ++n_fused;
      } else {
        // reductions over rank 1 of [n, m] operands
        uint32_t mask = 0;
        std::vector<std::pair<int, size_t>> refs;
        int64_t n0 = 0, m0 = 0;
        int rop = 0;
        for (int ps : prods) {
          const Step& r = steps[ps];
          const PNode& rn = nodes[r.out_node];
          if (r.ew || r.gemm_fused || r.group || !r.holder || rn.op != REDUCE_SUM || r.in_nodes.size() != 1 || rn.dtype != FLOAT) { ok = false; break; }
          uint32_t mk = 0;
          for (RankT q : eigen::unpack_rankset(*rn.func))
            if (q < rank_cap) mk |= 1u << q;
          const PNode& in = nodes[rn.args[0]];
          bool two_d = true;
          for (int q = 2; q < rank_cap; ++q)
            if (in.shape.at(q) != 1) two_d = false;
          if (!two_d || mk != 2u || in.dtype != FLOAT) { ok = false; break; }
          if (refs.empty()) { mask = mk; n0 = in.shape.at(0); m0 = in.shape.at(1); rop = rn.op; }
          else if ((int64_t)in.shape.at(0) != n0 || (int64_t)in.shape.at(1) != m0) { ok = false; break; }
          refs.push_back({r.in_nodes[0], r.in_offsets[0]});
        }
        if (!ok || refs.empty()) continue;
        // order by stack position when the operands are stacked already, else by graph position
        std::sort(refs.begin(), refs.end(), [&](const std::pair<int, size_t>& p, const std::pair<int, size_t>& q) {
          const PNode &a = nodes[p.first], &b = nodes[q.first];
          if (a.stack >= 0 && a.stack == b.stack) return a.stack_pos < b.stack_pos;
          return p < q;
        });
        const size_t slab = (size_t)n0 * (size_t)m0 * sizeof(float);
        int bn;
        size_t bo;
        if (!stack_operands(refs, slab, false, bn, bo)) continue;
        stack_operands(refs, slab, true, bn, bo);
        Step g;
        g.stack_reduce = true;
        g.pos = x.pos;
        g.out_node = x.out_node;
        g.red_op = rop;
        g.red_mask = mask;
        g.red_shape[0] = n0;
        g.red_shape[1] = m0 * T;
        g.in_nodes = {bn};
        g.in_offsets = {bo};
        for (auto& r : refs) g.dep_nodes.push_back(r.first);
        for (int ps : prods) { merge_acc(s, ps); drop_node(s, steps[ps].out_node); steps[ps].dead = true; }
        steps[s] = std::move(g);
        ++n_fused;
      }
    }

    // ================= (5) n-ary CONCAT of step outputs along the slowest rank -> produced in place =================
    for (int s = 0; s < ns; ++s) {
      Step& c = steps[s];
      if (c.dead || c.ew || c.gemm_fused || c.group || c.stack_reduce || c.kind != Step::NORMAL || !c.holder) continue;
      const PNode& cn = nodes[c.out_node];
      if (cn.op != CONCAT || c.in_nodes.size() < 3 || cn.exposed) continue;
      const int axis = (int)eigen::unpack_rank(*cn.func);
      bool ok = axis < rank_cap;
      for (int r = axis + 1; r < rank_cap && ok; ++r)
        if (cn.shape.at(r) != 1) ok = false;
      std::vector<std::pair<int, size_t>> refs;
      for (size_t i = 0; i < c.in_nodes.size() && ok; ++i) {
        const PNode& a = nodes[cn.args[i]];
        if (a.shape.at(axis) != 1 || a.dtype != cn.dtype) ok = false;
        refs.push_back({c.in_nodes[i], c.in_offsets[i]});
      }
      if (!ok) continue;
      const size_t slab = (size_t)nodes[refs[0].first].n * type_size(cn.dtype);
      int bn;
      size_t bo;
      bool same_root = true, fresh = true;
      for (auto& r : refs) {
        if (r.first != refs[0].first) same_root = false;
        if (nodes[r.first].stack >= 0) fresh = false;  // only fresh stacks: the CONCAT's result must begin at the stack's base
      }
      if (same_root || !fresh || !stack_operands(refs, slab, false, bn, bo)) continue;
      stack_operands(refs, slab, true, bn, bo);
      const int sid = nodes[refs[0].first].stack;
      nodes[c.out_node].alias_stack = sid;
      // readers of the CONCAT now wait for the producers of its operands
      for (int r : live_readers(c.out_node)) {
        drop_node(r, c.out_node);
        acc[r].rd.insert(acc[r].rd.end(), acc[s].rd.begin(), acc[s].rd.end());
        for (auto& m : refs) steps[r].dep_nodes.push_back(m.first);
      }
      c.dead = true;
      ++n_fused;
    }
    static const bool dbg = std::getenv("TCR_PLAN_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "fuse_recurrent_steps: %d rewrites over %d steps\n", n_fused, ns);
    if (n_fused == 0) { restore(); return; }

    // ---- compact + stable topological re-sort on (root, version)
    std::vector<int> live;
    for (int s = 0; s < ns; ++s)
      if (!steps[s].dead) live.push_back(s);
    const int nl = (int)live.size();
    std::map<std::pair<int, int>, int> writer;
    std::map<std::pair<int, int>, std::vector<int>> readers_at;
    for (int i = 0; i < nl; ++i) {
      for (auto& w : acc[live[i]].wr) writer[w] = i;
      for (auto& r : acc[live[i]].rd) readers_at[r].push_back(i);
    }
    std::vector<std::vector<int>> deps(nl);
    bool broken = false;
    for (int i = 0; i < nl && !broken; ++i) {
      for (auto& r : acc[live[i]].rd) {
        if (r.second == 0) continue;  // value from before the plan
        auto w = writer.find(r);
        if (w == writer.end()) {
          if (std::getenv("TCR_PLAN_DEBUG")) fprintf(stderr, "  step %s reads %s v%d (op %s) which nobody writes\n", step_name(steps[live[i]]).c_str(), nodes[r.first].shape.to_string().c_str(), r.second, nodes[r.first].func ? egen::name_op((_GENERATED_OPCODE)nodes[r.first].op).c_str() : "leaf");
          broken = true;
          break;
        }
        if (w->second != i) deps[i].push_back(w->second);
      }
      for (auto& w : acc[live[i]].wr) {
        if (w.second >= 2) {
          auto pw = writer.find({w.first, w.second - 1});
          if (pw != writer.end() && pw->second != i) deps[i].push_back(pw->second);
        }
        auto pr = readers_at.find({w.first, w.second - 1});
        if (pr != readers_at.end())
          for (int x : pr->second)
            if (x != i) deps[i].push_back(x);
      }
    }
    if (broken) { if (dbg) fprintf(stderr, "fuse_recurrent_steps: a read lost its writer, keeping the unfused plan\n"); restore(); return; }
    std::vector<int> indeg(nl, 0);
    std::vector<std::vector<int>> users(nl);
    for (int i = 0; i < nl; ++i) {
      std::sort(deps[i].begin(), deps[i].end());
      deps[i].erase(std::unique(deps[i].begin(), deps[i].end()), deps[i].end());
      for (int d : deps[i]) { users[d].push_back(i); ++indeg[i]; }
    }
    auto cmp = [&](int a, int b) { return steps[live[a]].pos > steps[live[b]].pos; };
    std::priority_queue<int, std::vector<int>, decltype(cmp)> ready(cmp);
    for (int i = 0; i < nl; ++i)
      if (indeg[i] == 0) ready.push(i);
    std::vector<int> order;
    while (!ready.empty()) {
      const int i = ready.top();
      ready.pop();
      order.push_back(i);
      for (int u : users[i])
        if (--indeg[u] == 0) ready.push(u);
    }
    if ((int)order.size() != nl && dbg) {
      // walk unsorted predecessors until a step repeats
      std::vector<int> seen(nl, 0);
      int cur = -1;
      for (int i = 0; i < nl; ++i)
        if (indeg[i] > 0) { cur = i; break; }
      for (int hop = 0; hop < 40 && cur >= 0; ++hop) {
        fprintf(stderr, "  [%d] %s %s\n", (int)steps[live[cur]].pos, step_name(steps[live[cur]]).c_str(), nodes[steps[live[cur]].out_node].shape.to_string().c_str());
        if (seen[cur]++) break;
        int nxt = -1;
        for (int d : deps[cur])
          if (indeg[d] > 0) { nxt = d; break; }
        cur = nxt;
      }
    }
    if ((int)order.size() != nl) { if (dbg) fprintf(stderr, "fuse_recurrent_steps: cycle (%d of %d sorted), keeping the unfused plan\n", (int)order.size(), nl); restore(); return; }  // a merge closed a cycle: keep the unfused plan
    std::vector<Step> sorted;
    sorted.reserve(nl);
    for (int i : order) sorted.push_back(std::move(steps[live[i]]));
    steps = std::move(sorted);
    for (auto& n : nodes) n.step = -1;
    std::vector<int> wr;
    for (size_t s = 0; s < steps.size(); ++s) {
      nodes[steps[s].out_node].step = (int)s;
      for (int e : steps[s].extra_outs) nodes[e].step = (int)s;
    }
  }

  // ---------------------------------------------------------------- elementwise steps over one iteration space -> one launch
  // After region building every elementwise result with more than one reader is its own launch. Backward through an unrolled
  // LSTM / GRU cell (tenncor/eteq/backprop.hpp:136-142 over cfg/tenncor/layer.yml:716-813) is, per time step, a chain of six such
  // results on the critical path (dh, dc, four gate pre-activation gradients), each ~5 us of launch latency for 256 KB of data.
  // Steps of equal iteration space that are adjacent in the (already topologically sorted) list are merged into one
  // tcr_elementwise_multi launch: programs run in list order per element, a later one reads an earlier one's result at the same
  // index. A step joins when (a) what it reads from the group is un-broadcast, whole and of the same size, and (b) no step lying
  // between the group's first member and it touches anything the group reads or writes — then every member can slide down to
  // the last member's position without crossing a dependency.
  // ---- symbolic form of a register-machine program (independent of the register allocation)
  struct Sym {
    int kind = 0;  // 0 input, 1 constant, 2 operation
    int op = 0, a = -1, b = -1, input = -1;
    double imm = 0;
  };
  static bool symbolic(const tcr_ew_program& q, std::vector<Sym>& ex, int& root) {
    int reg[NREGS];
    for (int r = 0; r < NREGS; ++r) reg[r] = -1;
    ex.clear();
    for (int k = 0; k < q.n_inputs; ++k) { Sym e; e.kind = 0; e.input = k; ex.push_back(e); reg[k] = k; }
    for (int i = 0; i < q.n_instrs; ++i) {
      const tcr_ew_instr& ins = q.instrs[i];
      Sym e;
      if (ins.op == TCR_EW_CONST) { e.kind = 1; e.imm = ins.imm; }
      else if (ins.op == TCR_EW_MOV) { if (reg[ins.a] < 0) return false; reg[ins.dst] = reg[ins.a]; continue; }
      else if (is_unary(ins.op)) { e.kind = 2; e.op = ins.op; e.a = reg[ins.a]; if (e.a < 0) return false; }
      else if (is_binary(ins.op) || ins.op == ADD || ins.op == MUL) { e.kind = 2; e.op = ins.op; e.a = reg[ins.a]; e.b = reg[ins.b]; if (e.a < 0 || e.b < 0) return false; }
      else return false;
      ex.push_back(e);
      reg[ins.dst] = (int)ex.size() - 1;
    }
    root = reg[q.outputs[0].reg];
    return root >= 0;
  }

  // Backward through one step of a gated cell as the derivative rules emit it (backprop.hpp:136-142):  s = a + b,
  // c = x*y + z*s, gates (x (1 - x)) * (y v) or (1 - x^2) * (y v) with v = s | c. Matched on expression trees modulo the
  // commutativity of ADD / MUL (bitwise commutative in IEEE arithmetic), so the hand-written kernel gives the same bits.
  bool match_cell_backward(Step& m) {
    const int count = (int)m.multi_progs.size();
    if (count < 3 || count > 8) return false;
    for (int k = 0; k < count; ++k) {
      const tcr_ew_program& q = m.multi_progs[k];
      if (q.dtype != FLOAT || q.outputs[0].dtype != FLOAT || q.n_outputs != 1) return false;
      for (auto& in : m.multi_inputs[k])
        if (in.mask != 0 || in.dtype != FLOAT || nodes[in.node].has_scalar) return false;
    }
    auto is_result_of = [&](int prog, int input, int producer) {
      const InputRef& r = m.multi_inputs[prog][input];
      return r.node == m.multi_outs[producer] && r.offset == 0;
    };
    std::vector<Sym> ex;
    int root = -1;
    auto leaf = [&](int e) { return e >= 0 && ex[e].kind == 0 ? ex[e].input : -1; };
    auto is_op = [&](int e, int op) { return e >= 0 && ex[e].kind == 2 && ex[e].op == op; };
    auto is_one = [&](int e) { return e >= 0 && ex[e].kind == 1 && ex[e].imm == 1.0; };
    // program 0: s = a + b
    if (!symbolic(m.multi_progs[0], ex, root) || !is_op(root, ADD) || leaf(ex[root].a) < 0 || leaf(ex[root].b) < 0 || leaf(ex[root].a) == leaf(ex[root].b)) return false;
    if (m.multi_progs[0].n_inputs != 2) return false;
    // program 1: c = x*y + z*s
    if (!symbolic(m.multi_progs[1], ex, root) || !is_op(root, ADD) || !is_op(ex[root].a, MUL) || !is_op(ex[root].b, MUL)) return false;
    {
      int l[4] = {leaf(ex[ex[root].a].a), leaf(ex[ex[root].a].b), leaf(ex[ex[root].b].a), leaf(ex[ex[root].b].b)};
      int s_at = -1;
      for (int k = 0; k < 4; ++k) {
        if (l[k] < 0) return false;
        if (is_result_of(1, l[k], 0)) { if (s_at >= 0) return false; s_at = k; }
      }
      if (s_at < 0) return false;
      const int z = l[s_at ^ 1], x = l[(s_at < 2) ? 2 : 0], y = l[(s_at < 2) ? 3 : 1];
      m.cb_c[0][0] = 1; m.cb_c[0][1] = x;
      m.cb_c[1][0] = 1; m.cb_c[1][1] = y;
      m.cb_c[2][0] = 1; m.cb_c[2][1] = z;
    }
    // gates
    for (int k = 2; k < count; ++k) {
      if (!symbolic(m.multi_progs[k], ex, root) || !is_op(root, MUL)) return false;
      int local = ex[root].a, up = ex[root].b;
      auto local_kind = [&](int e, int& x) {
        if (is_op(e, MUL)) {  // x * (1 - x)
          for (int sw = 0; sw < 2; ++sw) {
            const int xe = sw ? ex[e].b : ex[e].a, se = sw ? ex[e].a : ex[e].b;
            if (leaf(xe) >= 0 && is_op(se, SUB) && is_one(ex[se].a) && leaf(ex[se].b) == leaf(xe)) { x = leaf(xe); return 1; }
          }
        } else if (is_op(e, SUB) && is_one(ex[e].a) && is_op(ex[e].b, SQUARE) && leaf(ex[ex[e].b].a) >= 0) {  // 1 - x^2
          x = leaf(ex[ex[e].b].a);
          return 2;
        }
        return 0;
      };
      int x = -1;
      int kind = local_kind(local, x);
      if (!kind) { std::swap(local, up); kind = local_kind(local, x); }
      if (!kind || !is_op(up, MUL)) return false;
      int y = leaf(ex[up].a), v = leaf(ex[up].b);
      if (y < 0 || v < 0) return false;
      auto which = [&](int input) { return is_result_of(k, input, 0) ? 0 : is_result_of(k, input, 1) ? 1 : -1; };
      if (which(v) < 0) std::swap(y, v);
      if (which(v) < 0 || which(y) >= 0 || which(x) >= 0) return false;
      m.cb_kind[k - 2] = kind; m.cb_sel[k - 2] = which(v); m.cb_x[k - 2] = x; m.cb_y[k - 2] = y;
    }
    m.cell_bwd = true;
    return true;
  }

  void merge_elementwise_steps() {
    if (std::getenv("TCR_NO_EW_MERGE")) return;
    const int ns = (int)steps.size();
    if (ns < 2) return;
    constexpr int MAX_MEMBERS = 8, WINDOW = 12;
    auto candidate = [&](const Step& e) {
      return e.ew && !e.ew_reduce && !e.multi && !e.dead && e.kind == Step::NORMAL && e.prog.n_outputs == 1 && e.extra_outs.empty() && e.dep_nodes.empty() &&
             !is_assign(nodes[e.out_node].op) && nodes[e.out_node].offset == 0 && nodes[e.out_node].root == e.out_node &&
             e.prog.dims[0] * e.prog.dims[1] * e.prog.dims[2] < (1ll << 31);
    };
    std::vector<std::vector<int>> node_readers(nodes.size());  // steps reading each storage root
    {
      std::vector<int> tmp;
      for (int s = 0; s < ns; ++s) {
        if (steps[s].dead) continue;
        step_reads(steps[s], tmp);
        for (int x : tmp) node_readers[x].push_back(s);
      }
    }
    std::vector<int> group, between;
    std::vector<int> g_reads, g_writes;  // storage roots
    std::vector<int> rd, wr;
    int n_merged = 0;
    auto contains = [](const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };
    auto close = [&] {
      if (group.size() >= 2) {
        const int last = group.back();
        Step m = steps[last];
        m.multi = true;
        std::vector<InputRef> outside;
        for (int g : group) {
          const Step& e = steps[g];
          m.multi_progs.push_back(e.prog);
          m.multi_inputs.push_back(e.inputs);
          m.multi_outs.push_back(e.out_node);
          bool outside_reader = nodes[e.out_node].exposed;
          for (int r : node_readers[e.out_node])
            if (!contains(group, r)) outside_reader = true;
          m.multi_keep.push_back(outside_reader ? 1 : 0);
          for (auto& in : e.inputs)
            if (!contains(g_writes, in.node) && std::find(outside.begin(), outside.end(), in) == outside.end()) outside.push_back(in);
          if (g != last) { m.extra_outs.push_back(e.out_node); steps[g].dead = true; }
        }
        m.inputs = outside;
        if (!std::getenv("TCR_NO_CELL_BWD")) match_cell_backward(m);
        steps[last] = std::move(m);
        ++n_merged;
      }
      group.clear(); between.clear(); g_reads.clear(); g_writes.clear();
    };
    for (int s = 0; s < ns; ++s) {
      const Step& e = steps[s];
      if (e.dead) continue;
      if (!candidate(e)) {
        if (!group.empty()) {
          between.push_back(s);
          if ((int)between.size() > WINDOW) close();
        }
        continue;
      }
      bool join = !group.empty() && (int)group.size() < MAX_MEMBERS;
      if (join) {
        const Step& f = steps[group.front()];
        join = f.prog.dtype == e.prog.dtype && f.prog.dims[0] == e.prog.dims[0] && f.prog.dims[1] == e.prog.dims[1] && f.prog.dims[2] == e.prog.dims[2];
      }
      if (join)  // (a) reads from the group are whole and un-broadcast
        for (auto& in : e.inputs)
          if (contains(g_writes, in.node) && (in.mask != 0 || in.offset != 0 || nodes[in.node].n != nodes[e.out_node].n || nodes[in.node].dtype != nodes[e.out_node].dtype)) join = false;
      if (join) {  // (b) nothing in between touches the group's data, nor this step's
        step_reads(e, rd);
        for (int b : between) {
          std::vector<int> brd, bwr;
          step_reads(steps[b], brd);
          step_writes(steps[b], bwr);
          for (int x : brd) if (contains(g_writes, x)) join = false;
          for (int x : bwr) if (contains(g_writes, x) || contains(g_reads, x) || x == e.out_node) join = false;  // (this step itself stays behind them)
          if (steps[b].kind != Step::NORMAL) join = false;  // gradient exchange: keep its place
        }
      }
      if (!join) close();
      group.push_back(s);
      step_reads(e, rd);
      for (int x : rd) if (!contains(g_reads, x)) g_reads.push_back(x);
      g_writes.push_back(e.out_node);
      // steps skipped so far stay "between" the group's first member and whatever joins next
    }
    close();
    if (n_merged == 0) return;
    std::vector<Step> kept;
    kept.reserve(steps.size());
    for (auto& st : steps)
      if (!st.dead) kept.push_back(std::move(st));
    steps = std::move(kept);
    for (auto& n : nodes) n.step = -1;
    for (size_t k = 0; k < steps.size(); ++k) {
      if (steps[k].kind == Step::BUCKET_FLUSH) continue;
      nodes[steps[k].out_node].step = (int)k;
      for (int e : steps[k].extra_outs) nodes[e].step = (int)k;
    }
  }

  // ---------------------------------------------------------------- CONCAT steps -> one batched copy
  // An unrolled recurrent layer builds CONCAT(x_t, h_{t-1}) per time step (cfg/tenncor/layer.yml:716-813). The gate products read
  // x_t and h_{t-1} as K-segments, so those CONCATs are needed only by the weight-gradient product at the end of the backward
  // pass: every one of them slides down to the position of the last (same commutation rule as merge_elementwise_steps) and they
  // run as ONE tcr_copy2d_batched launch instead of T launches of ~4.5 us.
  void batch_concat_steps() {
    if (std::getenv("TCR_NO_CONCAT_BATCH")) return;
    const int ns = (int)steps.size();
    auto candidate = [&](const Step& e) {
      if (e.ew || e.dead || e.kind != Step::NORMAL || e.gemm_fused || e.conv_fused || e.group || e.stack_reduce || e.multi || !e.holder || !e.extra_outs.empty() ||
          !e.dep_nodes.empty())
        return false;
      const PNode& o = nodes[e.out_node];
      if (o.op != CONCAT || o.exposed || e.in_nodes.size() < 2 || o.root != e.out_node || o.offset != 0) return false;
      if (e.in_nodes.size() > 2)  // group concat: every operand has extent one along the axis (operator.hpp:336-368)
        for (int a : o.args)
          if (nodes[a].shape.at(eigen::unpack_rank(*o.func)) != 1) return false;
      return true;
    };
    auto contains = [](const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };
    std::vector<int> group, between, g_reads, g_writes, rd, brd, bwr;
    int n_merged = 0;
    auto close = [&] {
      if (group.size() >= 2) {
        const int last = group.back();
        Step m = steps[last];
        m.copy_batch = true;
        m.holder = nullptr;
        m.in_nodes.clear();
        m.in_offsets.clear();
        for (int g : group) {
          const Step& e = steps[g];
          const PNode& o = nodes[e.out_node];
          const int axis = eigen::unpack_rank(*o.func);
          const size_t es = type_size(o.dtype);
          int64_t inner = 1, outer = 1;
          for (int r = 0; r < axis; ++r) inner *= o.shape.at(r);
          for (int r = axis + 1; r < rank_cap; ++r) outer *= o.shape.at(r);
          const int64_t out_row = inner * o.shape.at(axis) * (int64_t)es;
          int64_t at = 0;
          for (size_t k = 0; k < e.in_nodes.size(); ++k) {
            const int64_t row = inner * nodes[o.args[k]].shape.at(axis) * (int64_t)es;
            tcr_copy2d_item it;
            it.dst = nullptr; it.src = nullptr;
            it.row_bytes = row; it.rows = outer; it.dst_pitch = out_row; it.src_pitch = row;
            m.copy_items.push_back(it);
            m.copy_dst.push_back(e.out_node); m.copy_dst_off.push_back((size_t)at);
            m.copy_src.push_back(e.in_nodes[k]); m.copy_src_off.push_back(e.in_offsets[k]);
            m.in_nodes.push_back(e.in_nodes[k]); m.in_offsets.push_back(e.in_offsets[k]);
            at += row;
          }
          if (g != last) { m.extra_outs.push_back(e.out_node); steps[g].dead = true; }
        }
        steps[last] = std::move(m);
        ++n_merged;
      }
      group.clear(); between.clear(); g_reads.clear(); g_writes.clear();
    };
    for (int s = 0; s < ns; ++s) {
      const Step& e = steps[s];
      if (e.dead) continue;
      if (!candidate(e)) { if (!group.empty()) between.push_back(s); continue; }
      bool join = !group.empty() && group.size() < 240;
      if (join) {
        for (int in : e.in_nodes) if (contains(g_writes, in)) join = false;  // a CONCAT of CONCATs keeps its own launch
        // the between-steps seen so far were checked when earlier members joined; the new ones since then:
        for (int b : between) {
          step_reads(steps[b], brd);
          step_writes(steps[b], bwr);
          for (int x : brd) if (contains(g_writes, x)) join = false;
          for (int x : bwr) if (contains(g_writes, x) || contains(g_reads, x)) join = false;
          if (steps[b].kind != Step::NORMAL) join = false;
        }
      }
      if (!join) close();
      else between.clear();  // verified against the group; later members add reads that these (earlier) steps cannot have written after them
      group.push_back(s);
      step_reads(e, rd);
      for (int x : rd) if (!contains(g_reads, x)) g_reads.push_back(x);
      g_writes.push_back(e.out_node);
    }
    close();
    if (n_merged == 0) return;
    std::vector<Step> kept;
    kept.reserve(steps.size());
    for (auto& st : steps)
      if (!st.dead) kept.push_back(std::move(st));
    steps = std::move(kept);
    for (auto& n : nodes) n.step = -1;
    for (size_t k = 0; k < steps.size(); ++k) {
      if (steps[k].kind == Step::BUCKET_FLUSH) continue;
      nodes[steps[k].out_node].step = (int)k;
      for (int e : steps[k].extra_outs) nodes[e].step = (int)k;
    }
  }

  // Reads / write of a step in terms of storage roots (before buffers exist)
  void step_access(const Step& st, std::vector<int>& reads, std::vector<int>& writes) const {
    step_reads(st, reads);
    step_writes(st, writes);
  }

  // Data-parallel plans: put every all-reduced gradient in one flat bucket and exchange it with ONE
  // collective. The steps are re-ordered by a stable topological sort in which every reader of a
  // reduced gradient waits for the flush, so all gradients are produced first, then the exchange,
  // then the optimiser updates and whatever reads the updated variables.
  void bucket_gradients() {
    if (std::getenv("TCR_NO_BUCKET")) return;
    std::vector<int> comm;
    for (size_t s = 0; s < steps.size(); ++s) {
      const Step& st = steps[s];
      if (!st.ew && !st.gemm_fused && nodes[st.out_node].op == IDENTITY && st.in_nodes.size() == 1) comm.push_back((int)s);
    }
    if (comm.size() < 2) return;
    const auto dt = nodes[steps[comm[0]].out_node].dtype;
    const double scale = dp::allreduce_scale(*nodes[steps[comm[0]].out_node].func);
    for (int c : comm) {
      const PNode& o = nodes[steps[c].out_node];
      if (o.dtype != dt || dp::allreduce_scale(*o.func) != scale || o.exposed || steps[c].in_offsets[0] != 0) return;
    }
    const int ns = (int)steps.size();
    // dependencies in the current order
    std::vector<std::vector<int>> deps(ns + 1);
    {
      std::unordered_map<int, int> last_writer;
      std::unordered_map<int, std::vector<int>> readers;
      std::vector<int> reads, writes;
      for (int s = 0; s < ns; ++s) {
        step_access(steps[s], reads, writes);
        for (int r : reads) {
          auto w = last_writer.find(r);
          if (w != last_writer.end()) deps[s].push_back(w->second);
          readers[r].push_back(s);
        }
        for (int write : writes) {
          auto w = last_writer.find(write);
          if (w != last_writer.end()) deps[s].push_back(w->second);
          auto rd = readers.find(write);
          if (rd != readers.end()) {
            for (int x : rd->second)
              if (x != s) deps[s].push_back(x);
            rd->second.clear();
          }
          last_writer[write] = s;
        }
      }
    }
    // the flush (index ns) waits for every member; whoever waited for a member waits for the flush
    std::set<int> is_comm(comm.begin(), comm.end());
    for (int s = 0; s < ns; ++s) {
      if (is_comm.count(s)) continue;
      bool reads_member = false;
      for (int d : deps[s]) reads_member |= is_comm.count(d) > 0;
      if (reads_member) deps[s].push_back(ns);
    }
    deps[ns] = comm;
    // stable topological order: smallest original index first; the flush sorts right after the last member
    std::vector<int> indeg(ns + 1, 0);
    std::vector<std::vector<int>> users(ns + 1);
    for (int s = 0; s <= ns; ++s) {
      std::sort(deps[s].begin(), deps[s].end());
      deps[s].erase(std::unique(deps[s].begin(), deps[s].end()), deps[s].end());
      for (int d : deps[s]) { users[d].push_back(s); ++indeg[s]; }
    }
    auto key = [&](int s) { return s == ns ? 2 * comm.back() + 1 : 2 * s; };
    auto cmp = [&](int a, int b) { return key(a) > key(b); };
    std::priority_queue<int, std::vector<int>, decltype(cmp)> ready(cmp);
    for (int s = 0; s <= ns; ++s)
      if (indeg[s] == 0) ready.push(s);
    std::vector<int> order;
    while (!ready.empty()) {
      int s = ready.top();
      ready.pop();
      order.push_back(s);
      for (int u : users[s])
        if (--indeg[u] == 0) ready.push(u);
    }
    if ((int)order.size() != ns + 1) return;  // cycle: keep the per-gradient exchange
    // slots
    size_t off = 0;
    std::vector<size_t> slot(ns, 0);
    for (int c : comm) {
      slot[c] = off;
      size_t bytes = (size_t)nodes[steps[c].out_node].n * type_size(dt);
      off += (bytes + 15) / 16 * 16;
    }
    bucket_bytes = off;
    bucket_dtype = dt;
    bucket_scale = scale;
    std::vector<Step> reordered;
    std::vector<int> new_index(ns + 1, -1);
    Step flush;
    flush.kind = Step::BUCKET_FLUSH;
    flush.out_node = steps[comm.back()].out_node;
    for (int s : order) {
      new_index[s] = (int)reordered.size();
      if (s == ns) { reordered.push_back(flush); continue; }
      Step st = std::move(steps[s]);
      if (is_comm.count(s)) {
        st.kind = Step::BUCKET_MEMBER;
        st.bucket_offset = slot[s];
      }
      reordered.push_back(std::move(st));
    }
    for (int c : comm) reordered[new_index[ns]].bucket_members.push_back(new_index[c]);
    steps = std::move(reordered);
    for (size_t s = 0; s < steps.size(); ++s)
      if (steps[s].kind != Step::BUCKET_FLUSH) nodes[steps[s].out_node].step = (int)s;
  }

  void* leaf_ptr(PNode& n) {
    void* p = n.tens->device().device_data();
    if (!p) global::fatalf("planner: %s has no data", n.tens->to_string().c_str());
    bound.push_back({(int)(&n - nodes.data()), p});
    return p;
  }

  void assign_buffers() {
    // last step that reads each base node
    std::vector<int> last_use(nodes.size(), -1);
    for (size_t s = 0; s < steps.size(); ++s) {
      Step& st = steps[s];
      if (st.ew) for (auto& in : st.inputs) last_use[in.node] = (int)s;
      else for (int in : st.in_nodes) last_use[in] = (int)s;
      for (int d : st.dep_nodes) last_use[d] = (int)s;
    }
    for (auto& n : nodes)
      if (!n.func) n.ptr = leaf_ptr(n);
    // stacks: operands that one fused step reads back to back live in ONE plan-owned buffer for the whole plan
    for (auto& stk : stacks) {
      stk.base = alloc_owned(stk.slab * stk.members.size());
      for (int m : stk.members) nodes[m].ptr = (char*)stk.base + (size_t)nodes[m].stack_pos * stk.slab;
    }
    for (auto& n : nodes)
      if (n.alias_stack >= 0) n.ptr = stacks[n.alias_stack].base;
    if (bucket_bytes > 0) {
      // inside the symmetric region when there is a peer path: the exchange is then one kernel over NVLink peer memory
      void* symm = nullptr;
      if (tcr_comm_symm_alloc(&symm, bucket_bytes) == TCR_OK && symm != nullptr) bucket = symm;
      else bucket = alloc_owned(bucket_bytes);
      check(tcr_memset(bucket, 0, bucket_bytes), "tcr_memset");  // the 16-byte padding between slots stays zero
      // a gradient read only by its exchange is produced straight into its slot
      for (size_t s = 0; s < steps.size(); ++s) {
        Step& st = steps[s];
        if (st.kind != Step::BUCKET_MEMBER) continue;
        PNode& in = nodes[st.in_nodes[0]];
        PNode& out = nodes[st.out_node];
        if (in.func && in.step >= 0 && !in.exposed && !is_assign(in.op) && last_use[st.in_nodes[0]] == (int)s && in.stack < 0 &&
            steps[in.step].out_node == st.in_nodes[0] && steps[in.step].kind == Step::NORMAL && in.n == out.n && in.dtype == out.dtype) {
          in.bucket_slot = (int64_t)st.bucket_offset;
          st.member_in_place = true;
        }
      }
    }
    std::multimap<size_t, void*> pool;  // plan-local free list: buffers are reused once their last reader has run
    std::vector<std::vector<int>> dying(steps.size());
    // A plan that updates no variable (inference, metrics) keeps every result in its own buffer, like the reference keeps every
    // functor's data (internal/eigen/device.hpp): a later evaluation re-runs only what a changed leaf reaches (stale_steps) and
    // reads everything else where it was left. Training plans are stale from top to bottom after every step (the ASSIGNs bump the
    // variables), so they recycle buffers by lifetime instead. TCR_PLAN_KEEP_MB bounds the memory spent on keeping (default 4096).
    bool keep_all = bucket_bytes == 0 && std::getenv("TCR_NO_PARTIAL") == nullptr;
    {
      size_t total = 0;
      for (auto& st : steps) {
        if (st.kind != Step::NORMAL || is_assign(nodes[st.out_node].op)) { keep_all = false; break; }
        total += (size_t)nodes[st.out_node].n * type_size(nodes[st.out_node].dtype);
        for (int e : st.extra_outs) total += (size_t)nodes[e].n * type_size(nodes[e].dtype);
      }
      const char* cap = std::getenv("TCR_PLAN_KEEP_MB");
      if (total > (size_t)(cap ? std::atoll(cap) : 4096) << 20) keep_all = false;
    }
    auto place_pooled = [&](int node, size_t s) {
      PNode& o = nodes[node];
      size_t bytes = (size_t)o.n * type_size(o.dtype);
      size_t bucket = bytes < 512 ? 512 : bytes;
      if (keep_all) { o.ptr = alloc_owned(bucket); return; }
      auto it = pool.lower_bound(bucket);
      if (it != pool.end() && it->first <= bucket * 2) {
        o.ptr = it->second;
        pool.erase(it);
      } else {
        o.ptr = alloc_owned(bucket);
      }
      int lu = last_use[node];
      if (lu >= (int)s) dying[lu].push_back(node);
      else dying[s].push_back(node);  // never read: reusable right after
    };
    for (size_t s = 0; s < steps.size(); ++s) {
      Step& st = steps[s];
      PNode& out = nodes[st.out_node];
      for (int e : st.extra_outs) {  // further results of a grouped launch
        PNode& eo = nodes[e];
        if (eo.stack >= 0) continue;
        if (eo.exposed) {
          auto op = dynamic_cast<DevOp*>(eo.holder);
          if (!op) global::fatalf("planner: target %s cannot hold data", eo.tens->to_string().c_str());
          eo.ptr = op->ensure_buffer(1, memory);
          bound.push_back({e, eo.ptr});
        } else {
          place_pooled(e, s);
        }
      }
      if (st.kind == Step::BUCKET_FLUSH) {
        continue;  // owns nothing
      } else if (out.stack >= 0 && st.kind == Step::NORMAL) {
        // placed above: slab of a stack
      } else if (st.kind == Step::BUCKET_MEMBER) {
        out.ptr = (char*)bucket + st.bucket_offset;  // lives for the whole plan, never pooled
      } else if (out.bucket_slot >= 0) {
        out.ptr = (char*)bucket + out.bucket_slot;
      } else if (st.conv_dimg && out.exposed) {
        // a target reads the slice through the CONV holder's own (full-size) buffer: only the slice is written
        auto op = dynamic_cast<DevOp*>(out.holder);
        if (!op) global::fatalf("planner: target %s cannot hold data", out.tens->to_string().c_str());
        out.ptr = op->ensure_buffer(1, memory);
        bound.push_back({st.out_node, out.ptr});
      } else if (st.conv_dimg) {
        // only one slice of this node's output is ever read (checked in fuse_conv_image_gradient): allocate that slice and
        // hand out the base it would have inside the full tensor. Plan-owned, never recycled.
        out.ptr = (char*)alloc_owned((size_t)convs[out.conv].out_n * sizeof(float)) - st.conv_out_offset;
      } else if (is_assign(out.op)) {
        out.ptr = nodes[nodes[out.args[0]].root].ptr;  // variable storage
        out.root = nodes[out.args[0]].root;
      } else if (out.exposed) {
        auto op = dynamic_cast<DevOp*>(out.holder);
        if (!op) global::fatalf("planner: target %s cannot hold data", out.tens->to_string().c_str());
        out.ptr = op->ensure_buffer(1, memory);
        bound.push_back({st.out_node, out.ptr});
      } else if (!st.ew && !st.gemm_fused && out.op == IDENTITY && st.in_nodes.size() == 1 && st.in_offsets[0] == 0 &&
                 nodes[st.in_nodes[0]].func && nodes[st.in_nodes[0]].step >= 0 && !nodes[st.in_nodes[0]].exposed &&
                 !is_assign(nodes[st.in_nodes[0]].op) && last_use[st.in_nodes[0]] == (int)s &&
                 nodes[st.in_nodes[0]].n == out.n && nodes[st.in_nodes[0]].dtype == out.dtype) {
        // gradient exchange in place: the all-reduce takes over its only producer's buffer (no copy)
        const int in = st.in_nodes[0];
        out.ptr = nodes[in].ptr;
        auto& dl = dying[s];
        dl.erase(std::remove(dl.begin(), dl.end(), in), dl.end());
        int lu = last_use[st.out_node];
        if (lu >= (int)s) dying[lu].push_back(st.out_node);
        else dying[s].push_back(st.out_node);
      } else {
        place_pooled(st.out_node, s);
      }
      for (int d : dying[s]) {
        PNode& dn = nodes[d];
        size_t bytes = (size_t)dn.n * type_size(dn.dtype);
        for (auto& b : owned)
          if (b.first == dn.ptr) { bytes = b.second; break; }
        pool.emplace(bytes, dn.ptr);
      }
    }
    // resolve pointers inside the steps
    for (auto& st : steps) {
      if (st.kind == Step::BUCKET_FLUSH) continue;
      PNode& out = nodes[st.out_node];
      if (st.multi) {
        for (size_t q = 0; q < st.multi_progs.size(); ++q) {
          for (size_t k = 0; k < st.multi_inputs[q].size(); ++k) {
            PNode& in = nodes[st.multi_inputs[q][k].node];
            if (!in.ptr) global::fatalf("planner: input %s of %s was never materialised", in.tens->to_string().c_str(), out.tens->to_string().c_str());
            st.multi_progs[q].inputs[k].ptr = (const char*)in.ptr + st.multi_inputs[q][k].offset;
          }
          st.multi_progs[q].outputs[0].ptr = nodes[st.multi_outs[q]].ptr;
        }
        st.out = out.ptr;
      } else if (st.ew) {
        for (size_t k = 0; k < st.inputs.size(); ++k) {
          PNode& in = nodes[st.inputs[k].node];
          if (!in.ptr) global::fatalf("planner: input %s of %s was never materialised", in.tens->to_string().c_str(), out.tens->to_string().c_str());
          st.prog.inputs[k].ptr = (const char*)in.ptr + st.inputs[k].offset;
        }
        st.prog.outputs[0].ptr = out.ptr;
        st.out = out.ptr;
      } else {
        st.out = out.ptr;
        if (st.conv_fused) st.conv_cols = alloc_owned((size_t)st.conv_rows * (size_t)st.conv_pitch * sizeof(float));
        for (size_t k = 0; k < st.in_nodes.size(); ++k) {
          PNode& in = nodes[st.in_nodes[k]];
          if (!in.ptr) global::fatalf("planner: input %s of %s was never materialised", in.tens->to_string().c_str(), out.tens->to_string().c_str());
          st.in.push_back((const char*)in.ptr + st.in_offsets[k]);
        }
        if (st.copy_batch)
          for (size_t k = 0; k < st.copy_items.size(); ++k) {
            st.copy_items[k].dst = (char*)nodes[st.copy_dst[k]].ptr + st.copy_dst_off[k];
            st.copy_items[k].src = (const char*)nodes[st.copy_src[k]].ptr + st.copy_src_off[k];
          }
        if (st.group) {
          tcr_gemm_group_desc& gd = st.gd;
          size_t k = 0;
          for (int q = 0; q < gd.segments; ++q) gd.a[q] = st.in[k++];
          for (int g = 0; g < gd.groups; ++g)
            for (int q = 0; q < gd.segments; ++q) gd.b[g][q] = st.in[k++];
          for (int g = 0; g < gd.groups; ++g) {
            gd.bias[g] = st.g_bias_slot[g] >= 0 ? st.in[st.g_bias_slot[g]] : nullptr;
            gd.out[g] = st.g_out_nodes[g] >= 0 ? nodes[st.g_out_nodes[g]].ptr : nullptr;
          }
          if (gd.cell) {
            gd.c_prev = st.g_cprev_slot >= 0 ? st.in[st.g_cprev_slot] : nullptr;
            gd.c_out = nodes[st.g_c_node].ptr;
            gd.h_out = nodes[st.g_h_node].ptr;
          }
        }
      }
    }
  }

  void launch_one(Step& st) {
    if (st.kind == Step::BUCKET_FLUSH) {
      check(tcr_allreduce_sum(bucket, (int64_t)(bucket_bytes / type_size((_GENERATED_DTYPE)bucket_dtype)), bucket_dtype, bucket_scale),
            "tcr_allreduce_sum");
    } else if (st.kind == Step::BUCKET_MEMBER) {
      if (!st.member_in_place) check(tcr_d2d(st.out, st.in[0], (size_t)nodes[st.out_node].n * type_size(nodes[st.out_node].dtype)), "tcr_d2d");
    } else if (st.copy_batch) {
      check(tcr_copy2d_batched(st.copy_items.data(), (int)st.copy_items.size()), "tcr_copy2d_batched");
    } else if (st.cell_bwd) {
      tcr_cell_backward_desc d;
      std::memset(&d, 0, sizeof(d));
      const int count = (int)st.multi_progs.size();
      d.n = nodes[st.out_node].n;
      d.s_a = st.multi_progs[0].inputs[0].ptr;
      d.s_b = st.multi_progs[0].inputs[1].ptr;
      d.c_x = st.multi_progs[st.cb_c[0][0]].inputs[st.cb_c[0][1]].ptr;
      d.c_y = st.multi_progs[st.cb_c[1][0]].inputs[st.cb_c[1][1]].ptr;
      d.c_z = st.multi_progs[st.cb_c[2][0]].inputs[st.cb_c[2][1]].ptr;
      d.s_out = st.multi_keep[0] ? st.multi_progs[0].outputs[0].ptr : nullptr;
      d.c_out = st.multi_keep[1] ? st.multi_progs[1].outputs[0].ptr : nullptr;
      d.n_gates = count - 2;
      for (int k = 0; k < count - 2; ++k) {
        d.kind[k] = st.cb_kind[k];
        d.sel[k] = st.cb_sel[k];
        d.x[k] = st.multi_progs[k + 2].inputs[st.cb_x[k]].ptr;
        d.y[k] = st.multi_progs[k + 2].inputs[st.cb_y[k]].ptr;
        d.out[k] = st.multi_progs[k + 2].outputs[0].ptr;
      }
      check(tcr_cell_backward(&d), "tcr_cell_backward");
    } else if (st.multi) check(tcr_elementwise_multi(st.multi_progs.data(), (int)st.multi_progs.size(), st.multi_keep.data()), "tcr_elementwise_multi");
    else if (st.ew_reduce) check(tcr_elementwise_reduce(&st.prog, st.out, st.red_post, st.red_imm), "tcr_elementwise_reduce");
    else if (st.ew) check(tcr_elementwise(&st.prog), "tcr_elementwise");
    else if (st.group) {
      tcr_gemm_group_desc d = st.gd;
      d.precision = gemm_precision();
      check(tcr_gemm_grouped(&d), "tcr_gemm_grouped");
    } else if (st.stack_reduce) {
      check(tcr_reduce(st.red_op, st.in[0], st.out, st.red_shape, st.red_mask, FLOAT), "tcr_reduce");
    } else if (st.conv_dimg) {
      tcr_gemm_desc d = st.gemm;
      d.precision = gemm_precision();
      check(tcr_gemm(st.in[0], st.in[1], st.conv_cols, &d), "tcr_gemm");
      check(tcr_col2im(st.conv_cols, (char*)st.out + st.conv_out_offset, st.conv_img, st.conv_win, st.conv_pitch, FLOAT), "tcr_col2im");
    } else if (st.conv_fused) {
      tcr_gemm_desc d = st.gemm;
      d.precision = gemm_precision();
      d.bias = st.in.size() > 2 ? st.in[2] : nullptr;
      // forward product: the kernel gathers the patch tiles itself when the view allows it (tcr_gemm_patches); the patch matrix is
      // written only for the shapes it declines (few channels, windows over other ranks) and for the kernel gradient
      static const bool implicit = std::getenv("TCR_NO_IMPLICIT_CONV") == nullptr;
      if (implicit && !st.conv_grad && st.conv_pitch == d.k) {
        const int rc = tcr_gemm_patches(st.in[0], st.in[1], st.out, &d, st.conv_img, st.conv_win);
        if (rc == TCR_OK) return;
        if (rc != TCR_ERR_UNSUPPORTED) check(rc, "tcr_gemm_patches");
      }
      check(tcr_im2col(st.in[0], st.conv_cols, st.conv_img, st.conv_win, st.conv_pitch, (int)sizeof(float)), "tcr_im2col");
      check(tcr_gemm(st.conv_cols, st.in[1], st.out, &d), "tcr_gemm");
    } else if (st.gemm_fused) {
      tcr_gemm_desc d = st.gemm;
      d.precision = d.dtype == FLOAT ? gemm_precision() : TCR_GEMM_EXACT;
      d.bias = (d.epilogue != TCR_EPI_NONE && st.in.size() > 2) ? st.in[2] : nullptr;
      d.aux = st.gemm_aux_slot >= 0 ? st.in[st.gemm_aux_slot] : nullptr;
      check(tcr_gemm(st.in[0], st.in[1], st.out, &d), "tcr_gemm");
    } else st.holder->launch_with(st.out, st.in);
  }

  void launch_steps() {
    for (auto& st : steps) launch_one(st);
  }

  // ---------------------------------------------------------------- capture lanes
  // Steps are ordered by the buffers they touch (read-after-write on producers, write-after-read
  // and write-after-write on recycled buffers and on variables updated in place). Independent
  // steps are captured on different lanes, so the graph runs them concurrently.
  std::vector<int> step_lane;
  std::vector<std::vector<int>> step_waits;  // cross-lane producers each step waits for
  std::vector<char> step_marks;              // step is awaited from another lane
  int lanes_used = 1;

  void schedule_lanes(int n_lanes) {
    const int ns = (int)steps.size();
    step_lane.assign(ns, 0);
    step_waits.assign(ns, {});
    step_marks.assign(ns, 0);
    lanes_used = 1;
    if (n_lanes <= 1 || ns < 3) return;
    std::unordered_map<const void*, int> last_writer;
    std::unordered_map<const void*, std::vector<int>> readers;
    std::vector<std::vector<int>> deps(ns);
    int last_comm = -1;
    for (int s = 0; s < ns; ++s) {
      Step& st = steps[s];
      std::vector<int>& d = deps[s];
      auto read = [&](const void* base) {
        auto w = last_writer.find(base);
        if (w != last_writer.end()) d.push_back(w->second);
        readers[base].push_back(s);
      };
      if (st.ew) for (auto& in : st.inputs) read(nodes[in.node].ptr);
      else for (int in : st.in_nodes) read(nodes[in].ptr);
      for (int dn : st.dep_nodes) read(nodes[dn].ptr);
      auto write = [&](const void* out) {
        auto w = last_writer.find(out);
        if (w != last_writer.end()) d.push_back(w->second);
        auto r = readers.find(out);
        if (r != readers.end()) {
          for (int x : r->second)
            if (x != s) d.push_back(x);
          r->second.clear();
        }
        last_writer[out] = s;
      };
      if (st.kind == Step::BUCKET_FLUSH) {  // reduces every slot of the bucket in place
        for (int m : st.bucket_members) write(nodes[steps[m].out_node].ptr);
      } else {
        write(nodes[st.out_node].ptr);
        for (int e : st.extra_outs) write(nodes[e].ptr);
      }
      // collectives (one communicator) and RAND_UNIF (one generator state) keep program order on lane 0
      const bool collective = st.kind == Step::BUCKET_FLUSH ||
                              (st.kind == Step::NORMAL && !st.ew && !st.gemm_fused && (nodes[st.out_node].op == IDENTITY || nodes[st.out_node].op == RAND_UNIF));
      if (collective) {
        if (last_comm >= 0) d.push_back(last_comm);
        last_comm = s;
      }
      std::sort(d.begin(), d.end());
      d.erase(std::unique(d.begin(), d.end()), d.end());
    }
    // A plan that is one long dependent chain (an unrolled recurrent layer: ~520 of C4's 569 steps) gains nothing from lanes and
    // pays for the cross-lane event nodes of the few steps beside it (C4: 6.89 ms on 4 lanes, 6.74 ms on one): with fewer than
    // 1.25 steps per step of the longest chain everything is captured on lane 0.
    {
      std::vector<int> depth(ns, 1);
      int longest = 1;
      for (int s = 0; s < ns; ++s) {
        for (int dstep : deps[s]) depth[s] = std::max(depth[s], depth[dstep] + 1);
        longest = std::max(longest, depth[s]);
      }
      if (!std::getenv("TCR_GRAPH_LANES") && ns >= 64 && (double)ns < 1.25 * (double)longest) return;
    }
    std::vector<int> lane_tail(n_lanes, -1);  // last step captured on each lane
    for (int s = 0; s < ns; ++s) {
      Step& st = steps[s];
      int lane = -1;
      if (st.kind == Step::BUCKET_FLUSH ||
          (st.kind == Step::NORMAL && !st.ew && !st.gemm_fused && (nodes[st.out_node].op == IDENTITY || nodes[st.out_node].op == RAND_UNIF)))
        lane = 0;
      if (lane < 0) {  // continue the lane of the latest producer when this step directly follows it there
        for (auto it = deps[s].rbegin(); it != deps[s].rend(); ++it)
          if (lane_tail[step_lane[*it]] == *it) { lane = step_lane[*it]; break; }
      }
      if (lane < 0) {  // otherwise the lane that has been idle longest
        lane = 0;
        for (int l = 1; l < n_lanes; ++l)
          if (lane_tail[l] < lane_tail[lane]) lane = l;
      }
      step_lane[s] = lane;
      lane_tail[lane] = s;
      if (lane + 1 > lanes_used) lanes_used = lane + 1;
      // one wait per foreign lane: its latest producer orders all earlier ones on that lane
      std::vector<int> latest(n_lanes, -1);
      for (int dstep : deps[s])
        if (step_lane[dstep] != lane && dstep > latest[step_lane[dstep]]) latest[step_lane[dstep]] = dstep;
      for (int l = 0; l < n_lanes; ++l)
        if (latest[l] >= 0) {
          step_waits[s].push_back(latest[l]);
          step_marks[latest[l]] = 1;
        }
    }
  }

  void capture_steps_on_lanes() {
    std::vector<int> mark(steps.size(), -1);
    int current = 0;
    for (size_t s = 0; s < steps.size(); ++s) {
      if (step_lane[s] != current) {
        check(tcr_graph_lane(step_lane[s]), "tcr_graph_lane");
        current = step_lane[s];
      }
      for (int w : step_waits[s]) check(tcr_graph_wait(mark[w]), "tcr_graph_wait");
      launch_one(steps[s]);
      if (step_marks[s]) check(tcr_graph_record(&mark[s]), "tcr_graph_record");
    }
    if (current != 0) check(tcr_graph_lane(0), "tcr_graph_lane");
  }

  void build(const TensSetT& targets, const TensSetT& ignored, eigen::RTMemptrT mem, bool lower_only = false) {
    memory = std::move(mem);
    precision = gemm_precision();
    std::vector<iTensor*> ordered(targets.begin(), targets.end());
    std::sort(ordered.begin(), ordered.end());
    for (auto t : ordered) {
      target_keys.push_back(t);
      target_refs.push_back(t->weak_from_this());
      collect(t, ignored);
    }
    resolve_views();
    for (auto t : ordered) nodes[nodes[index.at(t)].root].exposed = true;
    bool has_rand = false;
    for (auto& n : nodes) {
      if (!n.func) continue;
      if (!is_idempotent((_GENERATED_OPCODE)n.op)) always_run = true;
      if (n.op == RAND_UNIF) has_rand = true;
    }
    fuse();
    build_steps();
    fuse_recurrent_steps();
    merge_elementwise_steps();
    batch_concat_steps();
    bucket_gradients();
    if (lower_only) return;
    assign_buffers();
    prepare_partial();
    n_launch_steps = steps.size();
    // RAND_UNIF reads and advances the device-resident generator state: graph replays draw fresh numbers
    use_graph = std::getenv("TCR_NO_GRAPH") == nullptr && !steps.empty();
    uses_rand = has_rand;
    const char* lanes_env = std::getenv("TCR_GRAPH_LANES");
    int n_lanes = lanes_env ? std::atoi(lanes_env) : TCR_GRAPH_LANES;
    if (n_lanes > TCR_GRAPH_LANES) n_lanes = TCR_GRAPH_LANES;
    if (use_graph) schedule_lanes(n_lanes);
  }

  bool still_valid() const {
    if (precision != gemm_precision()) return false;
    for (size_t k = 0; k < target_refs.size(); ++k) {
      auto sp = target_refs[k].lock();
      if (!sp || sp.get() != target_keys[k]) return false;
    }
    for (int i : functor_order)
      if (&static_cast<eigen::iEigen&>(nodes[i].func->device()) != nodes[i].holder) return false;
    for (auto& b : bound)
      if (nodes[b.first].tens->device().device_data() != b.second) return false;
    return true;
  }

  // replicate the reference's version bookkeeping (functor.hpp:246-269, device.hpp:527-531)
  bool propagate_versions(size_t max_version) {
    bool any = false;
    for (int i : functor_order) {
      PNode& n = nodes[i];
      any |= static_cast<eigen::Observable*>(n.func)->prop_version(max_version);
      if (is_assign(n.op)) {
        auto target = static_cast<eigen::iMutableLeaf*>(nodes[nodes[n.args[0]].root].tens);
        target->upversion(nodes[n.args[1]].tens->get_meta().state_version() + 1);
        static_cast<eigen::iEigen&>(target->device()).mark_device_dirty();
      }
    }
    return any;
  }

  std::string step_name(const Step& st) const {
    if (st.kind == Step::BUCKET_FLUSH) return "ALLREDUCE bucket";
    const PNode& o = nodes[st.out_node];
    std::string what = egen::name_op((_GENERATED_OPCODE)o.op);
    if (st.kind == Step::BUCKET_MEMBER) return what + " -> bucket";
    if (st.copy_batch) return "CONCAT-BATCH(" + std::to_string(1 + st.extra_outs.size()) + ")";
    if (st.cell_bwd) return "CELL-BACKWARD(" + std::to_string(st.multi_progs.size() - 2) + " gates, " + std::to_string(st.inputs.size()) + " in)";
    if (st.multi) {
      int instrs = 0;
      for (auto& q : st.multi_progs) instrs += q.n_instrs;
      return what + " multi(" + std::to_string(st.multi_progs.size()) + " programs, " + std::to_string(instrs) + " instr, " + std::to_string(st.inputs.size()) + " in)";
    }
    if (st.ew_reduce) return std::string("SUM") + (st.red_post == 1 ? "/c" : st.red_post == 2 ? "*c" : "") + " of fused(" + std::to_string(st.prog.n_instrs) + " instr, " + std::to_string(st.prog.n_inputs) + " in)";
    if (st.ew) return what + " fused(" + std::to_string(st.prog.n_instrs) + " instr, " + std::to_string(st.prog.n_inputs) + " in)";
    if (st.group) {
      std::string k;
      for (int q = 0; q < st.gd.segments; ++q) k += (q ? "+" : "") + std::to_string(st.gd.seg_k[q]);
      return std::string(st.gd.b_trans ? "GEMM-SUM" : "GEMM-GROUP") + (st.gd.cell ? "+cell" : "") + " x" + std::to_string(st.gd.groups) + " m" + std::to_string(st.gd.m) + " n" +
             std::to_string(st.gd.n) + " k" + k;
    }
    if (st.stack_reduce) return what + "-STACK(" + std::to_string(st.dep_nodes.size()) + ")";
    const std::string mnk = " m" + std::to_string(st.gemm.m) + " n" + std::to_string(st.gemm.n) + " k" + std::to_string(st.gemm.k);
    if (st.conv_fused)
      return st.conv_dimg ? "CONV2D-dX GEMM+col2im" + mnk : std::string(st.conv_grad ? "CONV2D-dK" : "CONV2D") + " im2col+GEMM" + (st.gemm.epilogue ? "+bias" : "") + (st.gemm.activation ? "+act" : "") + mnk;
    if (st.gemm_fused && st.gemm.post_op) return std::string("GEMM*") + (st.gemm.post_op == TCR_POST_MUL_DSIGMOID ? "dsigmoid" : "dtanh") + mnk;
    if (st.gemm_fused && st.in_nodes.size() == 2) return "GEMM^T" + mnk;
    if (st.gemm_fused) return "GEMM+bias" + std::string(st.gemm.activation ? "+act" : "") + mnk;
    return what;
  }

  std::vector<StepTiming> profile(int repeats) {
    std::vector<StepTiming> out;
    void *e0 = nullptr, *e1 = nullptr;
    check(tcr_event_create(&e0), "tcr_event_create");
    check(tcr_event_create(&e1), "tcr_event_create");
    for (auto& st : steps) {
      PNode& o = nodes[st.out_node];
      StepTiming t;
      t.what = o.tens->to_string();
      if (st.multi || st.ew_reduce) t.what = step_name(st);
      else if (st.ew) t.what += " fused(" + std::to_string(st.prog.n_instrs) + " instr, " + std::to_string(st.prog.n_inputs) + " in)";
      t.shape = o.shape.to_string();
      t.bytes = (size_t)o.n * type_size(o.dtype);
      if (st.ew) {
        for (auto& in : st.inputs) {
          size_t n = 1;
          for (int r = 0; r < rank_cap; ++r)
            if (!((in.mask >> r) & 1u)) n *= o.shape.at(r);
          t.bytes += n * type_size(in.dtype);
        }
      } else {
        for (int in : st.in_nodes) t.bytes += (size_t)nodes[in].n * type_size(nodes[in].dtype);
        if (st.conv_dimg) t.what = "CONV2D-dX GEMM+col2im m" + std::to_string(st.gemm.m) + " n" + std::to_string(st.gemm.n) + " k" + std::to_string(st.gemm.k);
        else if (st.conv_fused) t.what = std::string(st.conv_grad ? "CONV2D-dK" : "CONV2D") + " im2col+GEMM" + (st.gemm.epilogue ? "+bias" : "") + (st.gemm.activation ? "+act" : "") +
                                    " m" + std::to_string(st.gemm.m) + " n" + std::to_string(st.gemm.n) + " k" + std::to_string(st.gemm.k);
        else if (st.gemm_fused && st.in.size() == 2) t.what = "GEMM^T m" + std::to_string(st.gemm.m) + " n" + std::to_string(st.gemm.n) + " k" + std::to_string(st.gemm.k);
        else if (st.gemm_fused) t.what = "GEMM+bias" + std::string(st.gemm.activation ? "+act" : "") + " m" + std::to_string(st.gemm.m) + " n" +
                                    std::to_string(st.gemm.n) + " k" + std::to_string(st.gemm.k);
        else if (o.op == CONTRACT || o.op == MATMUL || o.op == CONV)
          for (int a : o.args) t.what += " " + nodes[a].shape.to_string();
      }
      auto once = [&] { launch_one(st); };
      once();
      check(tcr_event_record(e0), "tcr_event_record");
      for (int r = 0; r < repeats; ++r) once();
      check(tcr_event_record(e1), "tcr_event_record");
      float ms = 0;
      check(tcr_event_elapsed_ms(e0, e1, &ms), "tcr_event_elapsed_ms");
      t.ms = ms / repeats;
      out.push_back(t);
    }
    tcr_event_destroy(e0);
    tcr_event_destroy(e1);
    return out;
  }

  // ---------------------------------------------------------------- re-evaluation of the stale part only
  // The reference recomputes a functor only when a child carries a newer version (tenncor/eteq/functor.hpp:246-269,
  // internal/eigen/device.hpp:555-570). The plan keeps, per functor, the version its own buffers hold (`seen_version`: versions are
  // global, another plan or the node evaluator may have bumped them without touching this plan's private buffers), and runs only
  // the steps that write a stale functor — plus the producers of operands whose buffer was recycled for another result since
  // (assign_buffers pools by lifetime), found by walking the steps backwards. When every step is needed the captured graph replays.
  std::vector<size_t> seen_version;     // per node; functors only
  std::vector<char> shared_buffer;      // per node: its buffer also holds another step's result at some point of a run
  std::vector<std::vector<int>> node_writers;  // per node: the steps that write it (a variable updated in place has several)
  size_t steps_run_last = 0;
  bool partial_ok = false;

  void prepare_partial() {
    seen_version.assign(nodes.size(), (size_t)-1);
    shared_buffer.assign(nodes.size(), 0);
    partial_ok = std::getenv("TCR_NO_PARTIAL") == nullptr;
    std::unordered_map<const void*, int> writers;
    std::vector<int> wr;
    node_writers.assign(nodes.size(), {});
    for (size_t k = 0; k < steps.size(); ++k) {
      const Step& st = steps[k];
      if (st.kind != Step::NORMAL) { partial_ok = false; continue; }  // gradient exchange: every rank runs the same launches
      // A plan that updates variables is stale from top to bottom after every run (the ASSIGNs bump what everything reads); the few
      // steps that are not (constant sub-expressions) are cheaper to replay inside the captured graph than to launch the rest eagerly
      if (is_assign(nodes[st.out_node].op)) partial_ok = false;
      step_writes(st, wr);
      for (int q : st.multi_outs) wr.push_back(q);
      std::sort(wr.begin(), wr.end());
      wr.erase(std::unique(wr.begin(), wr.end()), wr.end());
      std::set<const void*> ptrs;  // an ASSIGN names its variable's storage twice (result node and variable): one writer
      for (int w : wr) { ptrs.insert(nodes[w].ptr); node_writers[w].push_back((int)k); }
      for (const void* q : ptrs) ++writers[q];
    }
    for (auto& st : steps) {
      if (st.kind != Step::NORMAL) continue;
      step_writes(st, wr);
      for (int w : wr)
        if (writers[nodes[w].ptr] > 1 || nodes[w].stack >= 0 || nodes[w].alias_stack >= 0 || nodes[w].bucket_slot >= 0) shared_buffer[w] = 1;
    }
  }

  // steps to launch for this evaluation (size == steps.size() = everything); false = no functor of the plan is stale
  bool stale_steps(std::vector<int>& todo) {
    todo.clear();
    const int ns = (int)steps.size();
    std::vector<char> stale(nodes.size(), 0), needed(ns, 0);
    bool any = false;
    for (int i : functor_order)
      if (nodes[i].func->get_meta().state_version() != seen_version[i]) { stale[i] = 1; any = true; }
    if (!any) return false;
    std::vector<int> wr, rd;
    for (int s = 0; s < ns; ++s) {
      step_writes(steps[s], wr);
      for (int w : wr) if (stale[w]) needed[s] = 1;
      for (int q : steps[s].multi_outs) if (stale[q]) needed[s] = 1;
    }
    for (int s = ns - 1; s >= 0; --s) {
      if (!needed[s]) continue;
      step_reads(steps[s], rd);
      if (steps[s].multi)
        for (auto& ins : steps[s].multi_inputs) for (auto& in : ins) rd.push_back(in.node);
      for (int r0 : rd) {
        const int r = nodes[r0].root >= 0 ? nodes[r0].root : r0;  // a view reads its base
        if (!shared_buffer[r] && !shared_buffer[r0]) continue;
        for (int base : {r, r0})
          for (int producer : node_writers[base])
            if (producer < s) needed[producer] = 1;
      }
    }
    for (int s = 0; s < ns; ++s) if (needed[s]) todo.push_back(s);
    return true;
  }

  void mark_seen() {
    for (int i : functor_order) seen_version[i] = nodes[i].func->get_meta().state_version();
  }

  void run(size_t max_version) {
    bool changed = propagate_versions(max_version);
    steps_run_last = 0;
    std::vector<int> todo;
    const bool consider_partial = partial_ok && has_run && !always_run && max_version == std::numeric_limits<size_t>::max();
    if (consider_partial) {
      if (!stale_steps(todo)) return;
      if (todo.size() < steps.size()) {  // possibly none: only views of leaves are stale — their holders still learn about it below
        if (uses_rand) eteq::rng_flush();
        for (int s : todo) launch_one(steps[s]);
        for (auto& n : nodes)
          if (n.exposed && n.holder) n.holder->mark_device_dirty();
        mark_seen();
        steps_run_last = todo.size();
        if (!todo.empty()) ++g_stats.partial_runs;
        return;
      }
    } else if (has_run && !changed && !always_run) {
      return;
    }
    if (steps.empty()) { has_run = true; mark_seen(); return; }
    steps_run_last = steps.size();
    if (uses_rand) eteq::rng_flush();  // a seed() since the last run must reach the device before a replay
    if (use_graph && has_run) {
      if (!graph) {
        check(tcr_graph_begin(), "tcr_graph_begin");
        try {
          if (lanes_used > 1) capture_steps_on_lanes();
          else launch_steps();
        } catch (...) {
          void* dead = nullptr;
          tcr_graph_end(&dead);
          if (dead) tcr_graph_destroy(dead);
          throw;
        }
        check(tcr_graph_end(&graph), "tcr_graph_end");
      }
      check(tcr_graph_launch(graph), "tcr_graph_launch");
    } else {
      launch_steps();  // first run is eager: warms the arena and surfaces errors outside capture
    }
    for (auto& n : nodes)
      if (n.exposed && n.holder) n.holder->mark_device_dirty();
    has_run = true;
    mark_seen();
  }
};

struct PlanCache {
  struct Key {
    std::vector<iTensor*> t, ig;
    bool operator<(const Key& o) const { return t != o.t ? t < o.t : ig < o.ig; }
  };
  std::map<Key, std::unique_ptr<Plan>> plans;
};

static std::set<PlanEvaluator*>& live_evaluators() {
  static auto* s = new std::set<PlanEvaluator*>();
  return *s;
}

PlanEvaluator::PlanEvaluator() : cache_(new PlanCache()) { live_evaluators().insert(this); }
PlanEvaluator::~PlanEvaluator() {
  live_evaluators().erase(this);
  g_last_plan = nullptr;
}

void PlanEvaluator::drop_plans() {
  cache_->plans.clear();
  g_last_plan = nullptr;
}

void drop_all_plans() {
  for (auto e : live_evaluators()) e->drop_plans();
}

std::vector<std::string> describe_plan(const TensSetT& targets) {
  Plan plan;
  plan.build(targets, {}, nullptr, true);
  std::vector<std::string> out;
  const bool verbose = std::getenv("TCR_PLAN_VERBOSE") != nullptr;
  for (auto& st : plan.steps) {
    std::string line = plan.step_name(st) + " " + (st.kind == Step::BUCKET_FLUSH ? std::string() : plan.nodes[st.out_node].shape.to_string());
    if (verbose && st.kind != Step::BUCKET_FLUSH) {
      // storage roots read and written, and the program of an elementwise step: what a fusion pass would have to merge
      line += "  -> n" + std::to_string(st.out_node);
      for (int e : st.extra_outs) line += ",n" + std::to_string(e);
      line += "  <-";
      if (st.ew) {
        for (auto& in : st.inputs) line += " n" + std::to_string(in.node) + (in.offset ? "+" + std::to_string(in.offset) : "") + (in.mask ? "/b" + std::to_string(in.mask) : "");
        line += "  {";
        for (int k = 0; k < st.prog.n_instrs; ++k) {
          const tcr_ew_instr& ins = st.prog.instrs[k];
          line += " r" + std::to_string(ins.dst) + "=";
          if (ins.op == TCR_EW_CONST) line += std::to_string(ins.imm);
          else if (ins.op == TCR_EW_MOV) line += "r" + std::to_string(ins.a);
          else line += egen::name_op((_GENERATED_OPCODE)ins.op) + "(r" + std::to_string(ins.a) + ",r" + std::to_string(ins.b) + ")";
        }
        line += " } out r" + std::to_string(st.prog.outputs[0].reg);
      } else {
        for (size_t k = 0; k < st.in_nodes.size(); ++k) line += " n" + std::to_string(st.in_nodes[k]) + (st.in_offsets[k] ? "+" + std::to_string(st.in_offsets[k]) : "");
        for (int dnode : st.dep_nodes) line += " ~n" + std::to_string(dnode);
      }
    }
    out.push_back(line);
  }
  return out;
}

std::vector<StepTiming> profile_last_plan(int repeats) {
  if (!g_last_plan) global::fatal("profile_last_plan: no plan has been evaluated yet");
  return g_last_plan->profile(repeats);
}

void PlanEvaluator::evaluate(iDevice& device, const TensSetT& targets, const TensSetT& ignored) {
  auto cdev = dynamic_cast<Device*>(&device);
  if (!cdev) {  // a foreign device (mock, profiler): the reference traversal applies
    Evaluator fallback;
    fallback.evaluate(device, targets, ignored);
    return;
  }
  for (auto ig : ignored)
    if (nullptr != ig && nullptr == ig->device().device_data())
      global::throw_errf("cannot ignore tensor %s without existing data", ig->to_string().c_str());
  PlanCache::Key key;
  key.t.assign(targets.begin(), targets.end());
  key.ig.assign(ignored.begin(), ignored.end());
  std::sort(key.t.begin(), key.t.end());
  std::sort(key.ig.begin(), key.ig.end());
  auto it = cache_->plans.find(key);
  if (it != cache_->plans.end() && !it->second->still_valid()) {
    if (g_last_plan == it->second.get()) g_last_plan = nullptr;
    cache_->plans.erase(it);
    it = cache_->plans.end();
  }
  if (it == cache_->plans.end()) {
    if (cache_->plans.size() >= 64) {  // bounded: plans pin device buffers
      cache_->plans.clear();
      g_last_plan = nullptr;
    }
    ensure_device();
    std::unique_ptr<Plan> plan(new Plan());
    plan->build(targets, ignored, cdev->memory());
    it = cache_->plans.emplace(key, std::move(plan)).first;
  }
  Plan& plan = *it->second;
  plan.run(cdev->max_version_);
  g_last_plan = &plan;
  g_stats.nodes = plan.functor_order.size();
  g_stats.steps = plan.steps.size();
  g_stats.launches = plan.n_launch_steps;
  g_stats.graph = plan.graph != nullptr;
  g_stats.cached = cache_->plans.size();
  g_stats.steps_run = plan.steps_run_last;
}

}  // namespace cuda
