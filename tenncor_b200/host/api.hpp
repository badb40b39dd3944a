// api.hpp — the user API (TenncorAPI and its nn / layer / loss / approx / init / random
// sub-APIs). The reference generates these from cfg/tenncor/*.yml; every composite below
// expands to the same primitive opcodes in the same order (citations at each function in
// api.cpp), so the functor graphs the kernels see are the reference's graphs.
#ifndef TCR_HOST_API_HPP
#define TCR_HOST_API_HPP

#include <random>

#include "layr.hpp"

namespace tenncor {

using layr::ETensor;
using layr::ETensorsT;
using DimPairsT = std::pair<teq::DimT, teq::DimT>;

// ---- core (cfg/tenncor/core.yml)
ETensor cast(const ETensor& input, egen::_GENERATED_DTYPE dtype);
ETensor assign(const eteq::VarptrT& target, const ETensor& source);
ETensor assign_add(const eteq::VarptrT& target, const ETensor& source);
ETensor assign_sub(const eteq::VarptrT& target, const ETensor& source);
ETensor assign_mul(const eteq::VarptrT& target, const ETensor& source);
ETensor assign_div(const eteq::VarptrT& target, const ETensor& source);
ETensor identity(const ETensor& input, const ETensorsT& execute_in_parallel = {});
ETensor unary(egen::_GENERATED_OPCODE op, const ETensor& input);  // abs neg sin cos tan exp log sqrt round sigmoid tanh square cube
ETensor binary(egen::_GENERATED_OPCODE op, const ETensor& a, const ETensor& b);  // pow add sub mul div eq neq lt gt min max
ETensor binary(egen::_GENERATED_OPCODE op, const ETensor& a, double scalar);
ETensor binary(egen::_GENERATED_OPCODE op, double scalar, const ETensor& b);
inline ETensor abs(const ETensor& x) { return unary(egen::ABS, x); }
inline ETensor neg(const ETensor& x) { return unary(egen::NEG, x); }
inline ETensor sin(const ETensor& x) { return unary(egen::SIN, x); }
inline ETensor cos(const ETensor& x) { return unary(egen::COS, x); }
inline ETensor tan(const ETensor& x) { return unary(egen::TAN, x); }
inline ETensor exp(const ETensor& x) { return unary(egen::EXP, x); }
inline ETensor log(const ETensor& x) { return unary(egen::LOG, x); }
inline ETensor sqrt(const ETensor& x) { return unary(egen::SQRT, x); }
inline ETensor round(const ETensor& x) { return unary(egen::ROUND, x); }
inline ETensor sigmoid(const ETensor& x) { return unary(egen::SIGMOID, x); }
inline ETensor tanh(const ETensor& x) { return unary(egen::TANH, x); }
inline ETensor square(const ETensor& x) { return unary(egen::SQUARE, x); }
inline ETensor cube(const ETensor& x) { return unary(egen::CUBE, x); }
template <typename A, typename B> ETensor pow(const A& a, const B& b) { return binary(egen::POW, a, b); }
template <typename A, typename B> ETensor add(const A& a, const B& b) { return binary(egen::ADD, a, b); }
template <typename A, typename B> ETensor sub(const A& a, const B& b) { return binary(egen::SUB, a, b); }
template <typename A, typename B> ETensor mul(const A& a, const B& b) { return binary(egen::MUL, a, b); }
template <typename A, typename B> ETensor div(const A& a, const B& b) { return binary(egen::DIV, a, b); }
template <typename A, typename B> ETensor eq(const A& a, const B& b) { return binary(egen::EQ, a, b); }
template <typename A, typename B> ETensor neq(const A& a, const B& b) { return binary(egen::NEQ, a, b); }
template <typename A, typename B> ETensor lt(const A& a, const B& b) { return binary(egen::LT, a, b); }
template <typename A, typename B> ETensor gt(const A& a, const B& b) { return binary(egen::GT, a, b); }
template <typename A, typename B> ETensor min(const A& a, const B& b) { return binary(egen::MIN, a, b); }
template <typename A, typename B> ETensor max(const A& a, const B& b) { return binary(egen::MAX, a, b); }
ETensor min(const ETensorsT& args);
ETensor max(const ETensorsT& args);
ETensor if_then_else(const ETensor& condition, const ETensor& then, const ETensor& otherwise);
ETensor reverse(const ETensor& arg, const std::set<teq::RankT>& dims);
ETensor permute(const ETensor& arg, const teq::RanksT& order);
ETensor extend(const ETensor& arg, const teq::DimsT& bcast);
ETensor extend(const ETensor& arg, teq::RankT offset, const teq::DimsT& xlist);
ETensor extend_like(const ETensor& arg, const ETensor& like);
ETensor concat(const ETensor& left, const ETensor& right, teq::RankT axis);
ETensor concat(const ETensorsT& args, teq::RankT axis);
ETensor reshape(const ETensor& arg, teq::Shape shape);
ETensor reduce(egen::_GENERATED_OPCODE op, const ETensor& tens, std::set<teq::RankT> dims);
ETensor reduce(egen::_GENERATED_OPCODE op, const ETensor& tens, teq::RankT offset = 0, teq::RankT ndims = teq::rank_cap);
ETensor reduce_1d(egen::_GENERATED_OPCODE op, const ETensor& arg, teq::RankT dimension);
inline ETensor reduce_sum(const ETensor& t, teq::RankT offset = 0, teq::RankT ndims = teq::rank_cap) { return reduce(egen::REDUCE_SUM, t, offset, ndims); }
inline ETensor reduce_prod(const ETensor& t, teq::RankT offset = 0, teq::RankT ndims = teq::rank_cap) { return reduce(egen::REDUCE_PROD, t, offset, ndims); }
inline ETensor reduce_min(const ETensor& t, teq::RankT offset = 0, teq::RankT ndims = teq::rank_cap) { return reduce(egen::REDUCE_MIN, t, offset, ndims); }
inline ETensor reduce_max(const ETensor& t, teq::RankT offset = 0, teq::RankT ndims = teq::rank_cap) { return reduce(egen::REDUCE_MAX, t, offset, ndims); }
inline ETensor reduce_sum_1d(const ETensor& t, teq::RankT d) { return reduce_1d(egen::REDUCE_SUM, t, d); }
inline ETensor reduce_prod_1d(const ETensor& t, teq::RankT d) { return reduce_1d(egen::REDUCE_PROD, t, d); }
inline ETensor reduce_min_1d(const ETensor& t, teq::RankT d) { return reduce_1d(egen::REDUCE_MIN, t, d); }
inline ETensor reduce_max_1d(const ETensor& t, teq::RankT d) { return reduce_1d(egen::REDUCE_MAX, t, d); }
ETensor argmax(const ETensor& tens, teq::RankT return_dim = 8);
ETensor n_elems(const ETensor& arg);
ETensor n_dims(const ETensor& arg, teq::RankT rank);
ETensor slice(const ETensor& arg, eigen::PairVecT<teq::DimT> extents);
ETensor slice(const ETensor& arg, teq::DimT offset, teq::DimT extent, teq::RankT dimension);
ETensor pad(const ETensor& arg, eigen::PairVecT<teq::DimT> paddings);
ETensor pad(const ETensor& arg, const DimPairsT& padding, teq::RankT dimension);
ETensor stride(const ETensor& arg, const teq::DimsT& incrs);
ETensor scatter(const ETensor& arg, const teq::Shape& outshape, const teq::DimsT& incrs);
ETensor contract(const ETensor& a, const ETensor& b, eigen::PairVecT<teq::RankT> dims = {{0, 1}});
ETensor matmul(const ETensor& a, const ETensor& b);
ETensor convolution(const ETensor& image, const ETensor& kernel, const teq::RanksT& dims);
ETensor transpose(const ETensor& arg);
ETensor reduce_mean(const ETensor& arg);
ETensor reduce_mean_1d(const ETensor& arg, teq::RankT dimension);
ETensor reduce_variance(const ETensor& arg);
ETensor reduce_variance_1d(const ETensor& arg, teq::RankT dimension);
ETensor reduce_l2norm(const ETensor& arg, teq::RankT offset = 0, teq::RankT ndims = teq::rank_cap);
ETensor reduce_l2norm_1d(const ETensor& arg, teq::RankT dimension);
ETensor clip_by_range(const ETensor& arg, double minval, double maxval);
ETensor clip_by_l2norm(const ETensor& arg, double upper);
ETensor sum(const ETensorsT& args);
ETensor prod(const ETensorsT& args);
ETensor softmax(const ETensor& arg, teq::RankT offset = 0, teq::RankT ndims = teq::rank_cap);
ETensor relu(const ETensor& arg);
ETensor softplus(const ETensor& arg);
ETensor sign(const ETensor& x);

// ---- random (cfg/tenncor/random.yml)
namespace random {
ETensor rand_unif(const ETensor& a, const ETensor& b);
ETensor rand_binom_one(const ETensor& arg);
}  // namespace random

/// the host generator behind the initialisers and tc.unif_gen / tc.norm_gen (global::get_generator(), seeded by tenncor::seed)
std::mt19937_64& host_rng();

// ---- init (cfg/tenncor/init.yml); host RNG = std::mt19937_64 seeded by tenncor::seed
namespace init {
layr::InitF random_normal(double mean = 0, double stddev = 1, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
layr::InitF random_uniform(double minval = -0.05, double maxval = 0.05, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
layr::InitF zeros(egen::_GENERATED_DTYPE dtype = egen::default_dtype);
layr::InitF ones(egen::_GENERATED_DTYPE dtype = egen::default_dtype);
layr::InitF constants(double value, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
layr::InitF xavier_uniform(double factor = 1, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
layr::InitF xavier_normal(double factor = 1, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
inline layr::InitF glorot_uniform(double factor = 1, egen::_GENERATED_DTYPE dtype = egen::default_dtype) { return xavier_uniform(factor, dtype); }
inline layr::InitF glorot_normal(double factor = 1, egen::_GENERATED_DTYPE dtype = egen::default_dtype) { return xavier_normal(factor, dtype); }
/// normal values re-drawn (up to 5 times, then clipped) when further than 2 stddev from the mean (tenncor/layr/init.hpp:40-72)
layr::InitF truncated_normal(double mean = 0, double stddev = 1, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
/// `gain` on the diagonal of a 2-D shape (init.yml:170-195)
layr::InitF identity(double gain = 1, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
/// truncated normal with stddev = sqrt(factor / shape_factor(shape)); default shape factor = fanavg (init.yml:196-219)
layr::InitF variance_scaling(double factor, std::function<double(teq::Shape)> shape_factor = {}, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
}  // namespace init

// ---- nn (cfg/tenncor/nn.yml)
namespace nn {
ETensor fully_connect(const ETensorsT& lefts, const ETensorsT& rights, const ETensor& bias = nullptr,
                      eigen::PairVecT<teq::RankT> dims = {{0, 1}});
ETensor conv2d(const ETensor& image, const ETensor& kernel, const ETensor& bias = nullptr,
               const std::pair<DimPairsT, DimPairsT>& zero_paddings = {{0, 0}, {0, 0}});
ETensor dropout(const ETensor& input, const ETensor& drop_rate);
/// (input - mean) / sqrt(variance + eps) * scale + offset; mean / variance default to the whole-tensor statistics (nn.yml:128-208)
ETensor batch_normalization(const ETensor& input, ETensor offset, ETensor scale, ETensor eps, layr::UnaryF get_mean = {}, layr::UnaryF get_variance = {});
ETensor batch_normalization(const ETensor& input, double offset = 0, double scale = 1, double eps = -1, layr::UnaryF get_mean = {},
                            layr::UnaryF get_variance = {});
/// 2x2 mean / max over the two ranks `dims`, stride 2, built from STRIDE + SLICE (nn.yml:209-268)
ETensor mean_pool2d(const ETensor& arg, std::pair<teq::RankT, teq::RankT> dims = {0, 1});
ETensor max_pool2d(const ETensor& arg, std::pair<teq::RankT, teq::RankT> dims = {0, 1});
}  // namespace nn

// ---- layer (cfg/tenncor/layer.yml)
namespace layer {
ETensor bind(layr::UnaryF unary, const teq::Shape& inshape = teq::Shape(), egen::_GENERATED_DTYPE dtype = egen::default_dtype);
ETensor link(ETensorsT layers, ETensor input = nullptr);
ETensor dense(const ETensor& input, const ETensor& kernel, const ETensor& bias = nullptr, eigen::PairVecT<teq::RankT> dims = {{0, 1}});
ETensor dense(const ETensor& input, const teq::DimsT& hidden_dims, layr::InitF kernel_init = {}, layr::InitF bias_init = {},
              bool with_bias = true, const eigen::PairVecT<teq::RankT>& dims = {{0, 1}});
ETensor dense(const teq::Shape& inshape, const teq::DimsT& hidden_dims, layr::InitF kernel_init = {}, layr::InitF bias_init = {},
              bool with_bias = true, const eigen::PairVecT<teq::RankT>& dims = {{0, 1}}, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
ETensor conv2d(const ETensor& input, const ETensor& kernel, const ETensor& bias = nullptr,
               const std::pair<DimPairsT, DimPairsT>& zero_padding = {{0, 0}, {0, 0}});
ETensor conv2d(const DimPairsT& kernel_hw, teq::DimT in_ncol, teq::DimT out_ncol, layr::InitF kernel_init = {}, layr::InitF bias_init = {},
               const std::pair<DimPairsT, DimPairsT>& zero_padding = {{0, 0}, {0, 0}}, bool with_bias = true,
               egen::_GENERATED_DTYPE dtype = egen::default_dtype);
ETensor conv2d(const ETensor& input, teq::DimT out_ncol, const DimPairsT& kernel_hw, layr::InitF kernel_init, layr::InitF bias_init,
               const std::pair<DimPairsT, DimPairsT>& zero_padding, bool with_bias = true);
ETensor conv2d(const ETensor& input, teq::DimT out_ncol, const DimPairsT& kernel_hw, layr::InitF kernel_init, layr::InitF bias_init,
               const std::string& padding, bool with_bias = true);
ETensor rnn(const ETensor& input, const ETensor& init_state, const ETensor& cell, const layr::UnaryF& activation, teq::RankT seq_dim = 1);
ETensor rnn(teq::DimT indim, teq::DimT hidden_dim, const layr::UnaryF& activation, teq::DimT nseq, layr::InitF kernel_init = {},
            layr::InitF bias_init = {}, teq::RankT seq_dim = 1, bool with_bias = true, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
ETensor lstm(const ETensor& input, const ETensor& init_state, const ETensor& init_hidden, const ETensor& ggate, const ETensor& forgate,
             const ETensor& ingate, const ETensor& outgate, teq::RankT seq_dim = 1);
ETensor lstm(const teq::Shape& inshape, teq::DimT hidden_dim, teq::DimT nseq, layr::InitF kernel_init = {}, layr::InitF bias_init = {},
             teq::RankT seq_dim = 1, bool with_bias = true, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
ETensor gru(const ETensor& input, const ETensor& init_state, const ETensor& ugate, const ETensor& rgate, const ETensor& hgate,
            teq::RankT seq_dim = 1);
ETensor gru(const teq::Shape& inshape, teq::DimT hidden_dim, teq::DimT nseq, layr::InitF kernel_init = {}, layr::InitF bias_init = {},
            teq::RankT seq_dim = 1, bool with_bias = true, egen::_GENERATED_DTYPE dtype = egen::default_dtype);
/// nn.dropout, bypassed where `training` is 0 (layer.yml:441-478)
ETensor dropout(const ETensor& input, const ETensor& drop_rate, ETensor training = nullptr);
/// batch normalization with optional moving statistics: with `training` given, mean / variance are the batch statistics where
/// training != 0 and the momentum-updated moving statistics (ASSIGNed in place) elsewhere (layer.yml:479-633)
ETensor batch_normalization(ETensor input, ETensor offset, ETensor scale, ETensor eps, ETensor training = nullptr, ETensor momentum = nullptr,
                            layr::InitF moving_mean_init = {}, layr::InitF moving_var_init = {}, teq::RankT axis = teq::rank_cap);
layr::RBMLayer rbm(teq::DimT nvisible, teq::DimT nhidden, layr::InitF kernel_init = {}, layr::InitF bias_init = {}, bool with_bias = true,
                   egen::_GENERATED_DTYPE dtype = egen::default_dtype);
}  // namespace layer

// ---- loss (cfg/tenncor/loss.yml)
namespace loss {
ETensor sqr_diff(const ETensor& target, const ETensor& input);
ETensor mean_squared(const ETensor& target, const ETensor& input, teq::RankT axis = teq::rank_cap);
ETensor cross_entropy(const ETensor& target, const ETensor& input, float eps = std::numeric_limits<float>::epsilon());
}  // namespace loss

/// tcr::derive (tenncor/src/eteq.cpp:40-76). When a data-parallel group is active the
/// returned gradients are wrapped so that evaluation all-reduces them (see dp.hpp).
ETensorsT derive(const ETensor& root, const ETensorsT& targets);

// ---- approx (cfg/tenncor/approx.yml)
namespace approx {
layr::VarErrsT sgd(const ETensor& error, const eteq::VarptrsT& variables, double learning_rate = 0.5, layr::UnaryF apply = {});
layr::VarErrsT adagrad(const ETensor& error, const eteq::VarptrsT& variables, double learning_rate = 0.5,
                       double epsilon = std::numeric_limits<float>::epsilon(), layr::UnaryF apply = {});
layr::VarErrsT adam(const ETensor& error, const eteq::VarptrsT& variables, double step_rate = 0.001, double decay1 = 0.9,
                    double decay2 = 0.999, double epsilon = std::numeric_limits<float>::epsilon());
layr::VarErrsT adadelta(const ETensor& error, const eteq::VarptrsT& variables, double step_rate = 1, double decay = 0.9,
                        double offset = 0.0001, double epsilon = std::numeric_limits<float>::epsilon(), layr::UnaryF apply = {});
layr::VarErrsT rms_momentum(const ETensor& error, const eteq::VarptrsT& variables, double learning_rate = 0.5,
                            double discount_factor = 0.99, double epsilon = std::numeric_limits<float>::epsilon(), layr::UnaryF apply = {});
}  // namespace approx

void seed(uint64_t s);  // seeds RAND_UNIF's Philox key and the host initialiser RNG (tc.seed)

}  // namespace tenncor

namespace trainer {

/// node to evaluate once per training step: forward + backward + all assigns
/// (tenncor/trainer/apply_update.hpp:13-41)
layr::ETensor apply_update(const layr::ETensorsT& models, layr::ApproxF update, layr::ErrorF err_func);

/// contrastive-divergence RBM step (tenncor/trainer/rbm.hpp:42-165)
layr::VarErrsT bbernoulli_approx(const layr::VarErrsT& assocs, double learning_rate, double discount_factor);
layr::ETensor sample_v2h(const layr::RBMLayer& model, layr::ETensor vis);
layr::ETensor sample_h2v(const layr::RBMLayer& model, layr::ETensor hid);
layr::ETensor gibbs_hvh(const layr::RBMLayer& model, layr::ETensor hid);
/// contrastive-divergence chain: visible (given) -> hidden sample -> k-1 Gibbs steps -> visible / hidden means (rbm.hpp:65-81)
struct CDChainIO final {
  explicit CDChainIO(layr::ETensor visible, layr::ETensor hidden = nullptr) : visible_(std::move(visible)), hidden_(std::move(hidden)) {}
  layr::ETensor visible_, hidden_, visible_mean_, hidden_mean_;
};
/// CD-k statistics standing in for back-propagated gradients: (weight, <v h> - <v' h'>), (hbias, mean(h - h')), (vbias, mean(v - v'))
/// and, with a persistent chain, (persistent, next chain state) (rbm.hpp:83-146)
layr::VarErrsT cd_grad_approx(CDChainIO& io, const layr::RBMLayer& model, size_t cdk = 1, eteq::VarptrT persistent = nullptr);
using BErrorF = std::function<layr::ETensor(const layr::ETensor&, const layr::ETensor&)>;
layr::ETensor rbm(const layr::RBMLayer& model, layr::ETensor visible, double learning_rate, double discount_factor,
                  BErrorF err_func = {}, size_t cdk = 1);


/// Deep belief network: greedy layer-wise CD-k pre-training of a stack of RBMs, then a softmax (logistic) layer trained on the
/// top-level samples with a decaying learning rate (tenncor/trainer/dbn.hpp:13-229; demo/dbn_demo.py:60-118).
struct DBNTrainer final {
  DBNTrainer(const std::vector<layr::RBMLayer>& rbms, layr::ETensor dense, teq::RankT softmax_dim, teq::DimT batch_size, double pretrain_lr = 0.1,
             double train_lr = 0.1, size_t cdk = 10, double l2_reg = 0., double lr_scaling = 0.95);
  /// `train_in`: batch_size x input_size elements of `dtype`, row-major (the reference's ShapedArr)
  void pretrain(const void* train_in, egen::_GENERATED_DTYPE dtype, size_t nepochs = 100, std::function<void(size_t, size_t)> logger = {});
  void finetune(const void* train_in, const void* train_out, egen::_GENERATED_DTYPE dtype, size_t nepochs = 100, std::function<void(size_t)> logger = {});
  double reconstruction_cost(size_t layer);
  double training_cost();

  size_t nlayers_, input_size_, output_size_, batch_size_;
  eteq::VarptrT trainx_, trainy_;
  layr::ETensorsT sample_pipes_;
  std::vector<layr::ETensorsT> rupdates_;
  layr::ETensor tupdate_;
  layr::ETensorsT rcosts_;
  layr::ETensor tcost_;
};

}  // namespace trainer

#endif  // TCR_HOST_API_HPP
