// api.cpp — user API composites, initialisers, optimisers and trainers.
// Each function cites the yaml block of the reference it restates.
#include "api.hpp"

#include <cmath>
#include <random>

#include "dp.hpp"

namespace tenncor {

using namespace teq;
using namespace egen;
using eteq::make_functor;
using eteq::make_constant_like;
using eteq::VarptrT;

// ------------------------------------------------------------------ core.yml
ETensor cast(const ETensor& input, _GENERATED_DTYPE dtype) {  // core.yml:94-118
  marsh::Maps attrs;
  eigen::pack_attr(attrs, dtype);
  return eteq::make_tfuncattr(dtype, CAST, {input}, attrs);
}
ETensor assign(const VarptrT& t, const ETensor& s) { return make_functor(ASSIGN, {t, s}); }          // core.yml:119-128
ETensor assign_add(const VarptrT& t, const ETensor& s) { return make_functor(ASSIGN_ADD, {t, s}); }  // :129-138
ETensor assign_sub(const VarptrT& t, const ETensor& s) { return make_functor(ASSIGN_SUB, {t, s}); }  // :139-148
ETensor assign_mul(const VarptrT& t, const ETensor& s) { return make_functor(ASSIGN_MUL, {t, s}); }  // :149-158
ETensor assign_div(const VarptrT& t, const ETensor& s) { return make_functor(ASSIGN_DIV, {t, s}); }  // :159-170

ETensor identity(const ETensor& input, const ETensorsT& execute_in_parallel) {  // core.yml:171-187
  TensptrsT args = {input};
  args.insert(args.end(), execute_in_parallel.begin(), execute_in_parallel.end());
  return make_functor(IDENTITY, args);
}

ETensor unary(_GENERATED_OPCODE op, const ETensor& input) { return make_functor(op, {input}); }  // core.yml:188-278
ETensor binary(_GENERATED_OPCODE op, const ETensor& a, const ETensor& b) { return make_functor(op, {a, b}); }  // :279-633
ETensor binary(_GENERATED_OPCODE op, const ETensor& a, double scalar) { return make_functor(op, {a, make_constant_like(scalar, a)}); }
ETensor binary(_GENERATED_OPCODE op, double scalar, const ETensor& b) { return make_functor(op, {make_constant_like(scalar, b), b}); }

ETensor min(const ETensorsT& args) {  // core.yml:569-586
  if (args.empty()) global::fatal("cannot min without arguments");
  ETensor out = args[0];
  // binary() by name: an unqualified min(ETensor, ETensor) finds std::min through ADL (shared_ptr lives in std) and would compare pointers
  for (size_t i = 1, n = args.size(); i < n; ++i) out = binary(MIN, out, args[i]);
  return out;
}
ETensor max(const ETensorsT& args) {  // core.yml:616-633
  if (args.empty()) global::fatal("cannot max without arguments");
  ETensor out = args[0];
  for (size_t i = 1, n = args.size(); i < n; ++i) out = binary(MAX, out, args[i]);
  return out;
}

ETensor if_then_else(const ETensor& condition, const ETensor& then, const ETensor& otherwise) {  // core.yml:634-651
  if (then.get() == otherwise.get()) return then;
  return make_functor(SELECT, {condition, then, otherwise});
}
ETensor reverse(const ETensor& arg, const std::set<RankT>& dims) { return make_functor(REVERSE, {arg}, dims); }
ETensor permute(const ETensor& arg, const RanksT& order) { return make_functor(PERMUTE, {arg}, order); }
ETensor extend(const ETensor& arg, const DimsT& bcast) { return make_functor(EXTEND, {arg}, bcast); }
ETensor extend(const ETensor& arg, RankT offset, const DimsT& xlist) {  // core.yml:679-693
  DimsT bcast(offset, 1);
  bcast.insert(bcast.end(), xlist.begin(), xlist.end());
  return make_functor(EXTEND, {arg}, bcast);
}
ETensor extend_like(const ETensor& arg, const ETensor& like) { return make_functor(EXTEND, {arg}, like); }  // core.yml:714-723
ETensor concat(const ETensor& left, const ETensor& right, RankT axis) { return make_functor(CONCAT, {left, right}, axis); }
ETensor concat(const ETensorsT& args, RankT axis) { return make_functor(CONCAT, args, axis); }
ETensor reshape(const ETensor& arg, Shape shape) { return make_functor(RESHAPE, {arg}, shape); }

ETensor reduce(_GENERATED_OPCODE op, const ETensor& tens, std::set<RankT> dims) { return make_functor(op, {tens}, dims); }
ETensor reduce(_GENERATED_OPCODE op, const ETensor& tens, RankT offset, RankT ndims) {  // core.yml:773-872
  if (offset >= rank_cap) global::fatalf("cannot reduce dimensions [%d:]. Must be less than %d", (int)offset, (int)rank_cap);
  RanksT dims(std::min(ndims, (RankT)(rank_cap - offset)));
  std::iota(dims.begin(), dims.end(), offset);
  return make_functor(op, {tens}, std::set<RankT>(dims.begin(), dims.end()));
}
ETensor reduce_1d(_GENERATED_OPCODE op, const ETensor& arg, RankT dimension) {  // core.yml:1007-1082
  auto red = reduce(op, arg, dimension, 1);
  RanksT indices(rank_cap);
  auto bt = indices.begin();
  auto it = bt + dimension;
  std::iota(bt, it, 0);
  std::iota(it, indices.end(), dimension + 1);
  indices[rank_cap - 1] = dimension;
  return permute(red, indices);
}
ETensor argmax(const ETensor& tens, RankT return_dim) { return make_functor(ARGMAX, {tens}, return_dim); }
ETensor n_elems(const ETensor& arg) {  // core.yml:883-889: a scalar constant baked at build time
  return eteq::make_constant_scalar((double)arg->shape().n_elems(), Shape(), (_GENERATED_DTYPE)arg->get_meta().type_code());
}
ETensor n_dims(const ETensor& arg, RankT rank) {
  return eteq::make_constant_scalar((double)arg->shape().at(rank), Shape(), (_GENERATED_DTYPE)arg->get_meta().type_code());
}
ETensor slice(const ETensor& arg, eigen::PairVecT<DimT> extents) { return make_functor(SLICE, {arg}, extents); }
ETensor slice(const ETensor& arg, DimT offset, DimT extent, RankT dimension) {  // core.yml:918-927
  eigen::PairVecT<DimT> extents(std::max(rank_cap, dimension), {0, std::numeric_limits<DimT>::max()});
  extents[dimension] = {offset, extent};
  return slice(arg, extents);
}
ETensor pad(const ETensor& arg, eigen::PairVecT<DimT> paddings) { return make_functor(PAD, {arg}, paddings); }
ETensor pad(const ETensor& arg, const DimPairsT& padding, RankT dimension) {  // core.yml:937-952
  eigen::PairVecT<DimT> paddings(std::max(rank_cap, dimension), {0, 0});
  paddings[dimension] = padding;
  return pad(arg, paddings);
}
ETensor stride(const ETensor& arg, const DimsT& incrs) { return make_functor(STRIDE, {arg}, incrs); }
ETensor scatter(const ETensor& arg, const Shape& outshape, const DimsT& incrs) { return make_functor(SCATTER, {arg}, outshape, incrs); }
ETensor contract(const ETensor& a, const ETensor& b, eigen::PairVecT<RankT> dims) { return make_functor(CONTRACT, {a, b}, dims); }
ETensor matmul(const ETensor& a, const ETensor& b) { return make_functor(MATMUL, {a, b}); }
ETensor convolution(const ETensor& image, const ETensor& kernel, const RanksT& dims) { return make_functor(CONV, {image, kernel}, dims); }
ETensor transpose(const ETensor& arg) { return permute(arg, {1, 0}); }
ETensor reduce_mean(const ETensor& arg) { return div(reduce_sum(arg), n_elems(arg)); }  // core.yml:1090-1096
ETensor reduce_mean_1d(const ETensor& arg, RankT dimension) {                            // core.yml:1097-1109
  auto red = reduce_sum_1d(arg, dimension);
  auto dim = make_constant_like((double)arg->shape().at(dimension), red);
  return div(red, dim);
}
ETensor reduce_variance(const ETensor& arg) { return reduce_mean(square(sub(arg, extend_like(reduce_mean(arg), arg)))); }
ETensor reduce_variance_1d(const ETensor& arg, RankT dimension) {
  return reduce_mean_1d(square(sub(arg, extend_like(reduce_mean_1d(arg, dimension), arg))), dimension);
}
ETensor reduce_l2norm(const ETensor& arg, RankT offset, RankT ndims) { return sqrt(reduce_sum(square(arg), offset, ndims)); }
ETensor reduce_l2norm_1d(const ETensor& arg, RankT dimension) { return sqrt(reduce_sum_1d(square(arg), dimension)); }
ETensor clip_by_range(const ETensor& arg, double minval, double maxval) {  // core.yml:1147-1166
  if (minval > maxval) global::fatal("min value is below max");
  auto lo = make_constant_like(minval, arg), hi = make_constant_like(maxval, arg);
  return binary(MAX, binary(MIN, arg, hi), lo);  // not min()/max(): ADL would pick std::min / std::max on the shared_ptrs
}
ETensor clip_by_l2norm(const ETensor& arg, double upper) {  // core.yml:1167-1189
  if (upper == 0) global::fatal("cannot clip_by_norm with a upper limit of 0");
  auto norm = extend_like(reduce_l2norm(arg), arg);
  auto limit = make_constant_like(upper, arg);
  return if_then_else(lt(norm, limit), arg, div(mul(arg, limit), norm));
}
ETensor sum(const ETensorsT& args) {  // core.yml:1190-1211
  switch (args.size()) {
    case 0: global::fatal("cannot sum without arguments");
    case 1: return args[0];
    case 2: return add(args[0], args[1]);
    default: break;
  }
  return make_functor(ADD, args);
}
ETensor prod(const ETensorsT& args) {  // core.yml:1212-1231
  switch (args.size()) {
    case 0: global::fatal("cannot prod without arguments");
    case 1: return args[0];
    default: break;
  }
  return make_functor(MUL, args);
}
ETensor softmax(const ETensor& arg, RankT offset, RankT ndims) {  // core.yml:1232-1260
  if (offset + ndims > rank_cap) global::fatalf("cannot perform softmax on dimensions beyond %d", (int)rank_cap);
  auto overflow_preventer = extend_like(reduce_max(arg, offset, ndims), arg);
  auto exarg = exp(sub(arg, overflow_preventer));
  return div(exarg, extend_like(add(reduce_sum(exarg, offset, ndims), (double)std::numeric_limits<float>::epsilon()), exarg));
}
ETensor relu(const ETensor& arg) { return max(arg, 0.0); }
ETensor softplus(const ETensor& arg) { return log(add(1.0, exp(arg))); }
ETensor sign(const ETensor& x) { return add(mul(-2.0, lt(x, 0.0)), 1.0); }

// ------------------------------------------------------------------ random.yml
namespace random {
ETensor rand_unif(const ETensor& a, const ETensor& b) { return make_functor(RAND_UNIF, {a, b}); }
ETensor rand_binom_one(const ETensor& arg) {  // random.yml:22-32
  auto dtype = (_GENERATED_DTYPE)arg->get_meta().type_code();
  auto trial = rand_unif(eteq::make_variable_scalar(0, arg->shape(), "0", dtype), eteq::make_variable_scalar(1, arg->shape(), "1", dtype));
  return lt(trial, arg);
}
}  // namespace random

// ------------------------------------------------------------------ init.yml
std::mt19937_64& host_rng() {
  static std::mt19937_64 rng(0);
  return rng;
}

void seed(uint64_t s) {
  eteq::seed(s);
  host_rng().seed(s);
}

static double fanio(Shape shape) {  // layr::fanio (tenncor/layr/init.hpp:24-29)
  auto slist = narrow_shape(shape);
  return (double)std::accumulate(slist.begin(), slist.end(), (DimT)0);
}

static VarptrT var_from(const std::vector<double>& vals, _GENERATED_DTYPE dtype, Shape shape, const std::string& label) {
  std::vector<char> buf(vals.size() * type_size(dtype));
  type_convert(buf.data(), dtype, vals.data(), DOUBLE, vals.size());
  return eteq::make_variable(buf.data(), dtype, shape, label);
}

namespace init {
layr::InitF random_normal(double mean, double stddev, _GENERATED_DTYPE dtype) {
  return [=](Shape shape, std::string label) {
    std::normal_distribution<double> dist(mean, stddev);
    std::vector<double> vec(shape.n_elems());
    for (auto& v : vec) v = dist(host_rng());
    return var_from(vec, dtype, shape, label);
  };
}
layr::InitF random_uniform(double minval, double maxval, _GENERATED_DTYPE dtype) {
  return [=](Shape shape, std::string label) {
    std::uniform_real_distribution<double> dist(minval, maxval);
    std::vector<double> vec(shape.n_elems());
    for (auto& v : vec) v = dist(host_rng());
    return var_from(vec, dtype, shape, label);
  };
}
layr::InitF constants(double value, _GENERATED_DTYPE dtype) {
  return [=](Shape shape, std::string label) { return eteq::make_variable_scalar(value, shape, label, dtype); };
}
layr::InitF zeros(_GENERATED_DTYPE dtype) { return constants(0, dtype); }
layr::InitF ones(_GENERATED_DTYPE dtype) { return constants(1, dtype); }
layr::InitF xavier_uniform(double factor, _GENERATED_DTYPE dtype) {  // init.yml:147-169
  return [=](Shape shape, std::string label) {
    double bound = factor * std::sqrt(6. / fanio(shape));
    std::uniform_real_distribution<double> dist(-bound, bound);
    std::vector<double> vec(shape.n_elems());
    for (auto& v : vec) v = dist(host_rng());
    return var_from(vec, dtype, shape, label);
  };
}
layr::InitF xavier_normal(double factor, _GENERATED_DTYPE dtype) {  // init.yml:115-136 (truncated at 2 sigma)
  return [=](Shape shape, std::string label) {
    double stdev = factor * std::sqrt(2. / fanio(shape));
    std::normal_distribution<double> dist(0, stdev);
    std::vector<double> vec(shape.n_elems());
    for (auto& v : vec) {
      v = dist(host_rng());
      for (int retry = 0; std::abs(v) > 2 * stdev && retry < 5; ++retry) v = dist(host_rng());
    }
    return var_from(vec, dtype, shape, label);
  };
}
static void truncated_fill(std::vector<double>& vec, double mean, double stdev) {  // layr::truncated_normal, init.hpp:40-72
  std::normal_distribution<double> dist(mean, stdev);
  const double upper = mean + 2 * stdev, lower = mean - 2 * stdev;
  for (auto& v : vec) {
    v = dist(host_rng());
    for (int retry = 0; (v > upper || v < lower) && retry < 5; ++retry) v = dist(host_rng());
    v = std::min(std::max(v, lower), upper);
  }
}
layr::InitF truncated_normal(double mean, double stddev, _GENERATED_DTYPE dtype) {  // init.yml:54-74
  return [=](Shape shape, std::string label) {
    std::vector<double> vec(shape.n_elems());
    truncated_fill(vec, mean, stddev);
    return var_from(vec, dtype, shape, label);
  };
}
layr::InitF identity(double gain, _GENERATED_DTYPE dtype) {  // init.yml:170-195
  return [=](Shape shape, std::string label) {
    if (false == shape.compatible_after(Shape(), 2)) global::fatal("identity initialization can only be used for to 2D tensors");
    std::vector<double> vec(shape.n_elems(), 0);
    const DimT x = shape.at(0), y = shape.at(1);
    for (DimT diag = 0, n = std::min(x, y); diag < n; ++diag) vec[diag + diag * x] = gain;
    return var_from(vec, dtype, shape, label);
  };
}
layr::InitF variance_scaling(double factor, std::function<double(Shape)> shape_factor, _GENERATED_DTYPE dtype) {  // init.yml:196-219
  if (!shape_factor) shape_factor = [](Shape shape) { return fanio(shape) / 2; };  // layr::fanavg
  return [=](Shape shape, std::string label) {
    std::vector<double> vec(shape.n_elems());
    truncated_fill(vec, 0, std::sqrt(factor / shape_factor(shape)));
    return var_from(vec, dtype, shape, label);
  };
}
}  // namespace init

// ------------------------------------------------------------------ nn.yml
namespace nn {
ETensor fully_connect(const ETensorsT& lefts, const ETensorsT& rights, const ETensor& bias, eigen::PairVecT<RankT> dims) {  // nn.yml:14-47
  size_t nlefts = lefts.size();
  if (nlefts != rights.size())
    global::fatalf("number of lefts (%d) must equal the number of rights (%d)", (int)nlefts, (int)rights.size());
  auto out = contract(lefts[0], rights[0], dims);
  for (size_t i = 1; i < nlefts; ++i) out = add(out, contract(lefts[i], rights[i], dims));
  if (nullptr != bias) out = add(out, extend_like(bias, out));
  return out;
}

ETensor conv2d(const ETensor& image, const ETensor& kernel, const ETensor& bias, const std::pair<DimPairsT, DimPairsT>& zp) {  // nn.yml:48-98
  ETensor cimage = image;
  if (zp.first.first > 0 || zp.first.second > 0 || zp.second.first > 0 || zp.second.second > 0)
    cimage = pad(cimage, eigen::PairVecT<DimT>{{0, 0}, {zp.first.first, zp.first.second}, {zp.second.first, zp.second.second}});
  DimT img_pad = kernel->shape().at(0) - 1;  // out
  cimage = pad(cimage, DimPairsT{img_pad, img_pad}, 4);
  auto out = permute(convolution(cimage, reverse(kernel, {0}), {4, 0, 1, 2}), {4, 1, 2, 3});
  if (nullptr != bias) out = add(out, extend_like(bias, out));
  return out;
}

ETensor dropout(const ETensor& input, const ETensor& drop_rate) {  // nn.yml:112-130
  ETensor rate = sub(make_constant_like(1, drop_rate), drop_rate);
  if (false == rate->shape().compatible_after(input->shape(), 0)) rate = extend_like(rate, input);
  auto mask = random::rand_binom_one(rate);
  auto denom = div(reduce_sum(mask), n_elems(mask));
  return mul(input, div(mask, extend_like(denom, mask)));
}

ETensor batch_normalization(const ETensor& input, ETensor offset, ETensor scale, ETensor eps, layr::UnaryF get_mean, layr::UnaryF get_variance) {  // nn.yml:158-208
  if (!get_mean) get_mean = [](const ETensor& in) { return extend_like(reduce_mean(in), in); };
  if (!get_variance) get_variance = [](const ETensor& in) { return extend_like(reduce_variance(in), in); };
  if (false == offset->shape().compatible_after(input->shape(), 0)) offset = extend_like(offset, input);
  if (false == scale->shape().compatible_after(input->shape(), 0)) scale = extend_like(scale, input);
  if (false == eps->shape().compatible_after(input->shape(), 0)) eps = extend_like(eps, input);
  auto norm = div(sub(input, get_mean(input)), sqrt(add(get_variance(input), eps)));
  return add(mul(norm, scale), offset);
}

static double type_epsilon(const ETensor& t) {  // std::numeric_limits<T>::epsilon() of the input's element type
  return (_GENERATED_DTYPE)t->get_meta().type_code() == DOUBLE ? std::numeric_limits<double>::epsilon() : std::numeric_limits<float>::epsilon();
}

ETensor batch_normalization(const ETensor& input, double offset, double scale, double eps, layr::UnaryF get_mean, layr::UnaryF get_variance) {  // nn.yml:128-157
  if (eps < 0) eps = type_epsilon(input);
  return batch_normalization(input, eteq::make_constant_like(offset, input), eteq::make_constant_like(scale, input),
                             eteq::make_constant_like(eps, input), get_mean, get_variance);
}

static ETensor pool2d(const ETensor& arg, std::pair<RankT, RankT> dims, bool take_max) {  // nn.yml:209-268
  Shape shape = arg->shape();
  DimT xextent = shape.at(dims.first) - 1, yextent = shape.at(dims.second) - 1;
  DimsT strider(rank_cap, 1);
  strider[dims.first] = strider[dims.second] = 2;
  auto top_left = stride(arg, strider);
  auto top_right = stride(slice(arg, 1, xextent, dims.first), strider);
  auto bot_left = stride(slice(arg, 1, yextent, dims.second), strider);
  eigen::PairVecT<DimT> pvec(rank_cap, {0, std::numeric_limits<DimT>::max()});
  pvec[dims.first] = {1, xextent};
  pvec[dims.second] = {1, yextent};
  auto bot_right = stride(slice(arg, pvec), strider);
  ETensorsT corners = {top_left, top_right, bot_left, bot_right};
  return take_max ? max(corners) : div(sum(corners), 4.);
}
ETensor mean_pool2d(const ETensor& arg, std::pair<RankT, RankT> dims) { return pool2d(arg, dims, false); }
ETensor max_pool2d(const ETensor& arg, std::pair<RankT, RankT> dims) { return pool2d(arg, dims, true); }
}  // namespace nn

// ------------------------------------------------------------------ layer.yml
namespace layer {

ETensor dropout(const ETensor& input, const ETensor& drop_rate, ETensor training) {  // layer.yml:455-478
  auto out = nn::dropout(input, drop_rate);
  if (nullptr != training) {
    if (false == training->shape().compatible_after(input->shape(), 0)) training = extend_like(training, input);
    out = if_then_else(training, out, input);
  }
  return out;
}

ETensor batch_normalization(ETensor input, ETensor offset, ETensor scale, ETensor eps, ETensor training, ETensor momentum,
                            layr::InitF moving_mean_init, layr::InitF moving_var_init, RankT axis) {  // layer.yml:523-633
  layr::UnaryF get_mean, get_var;
  const bool whole = axis >= rank_cap;
  auto batch_mean = [whole, axis](const ETensor& in) { return extend_like(whole ? reduce_mean(in) : reduce_mean_1d(in, axis), in); };
  auto batch_var = [whole, axis](const ETensor& in) { return extend_like(whole ? reduce_variance(in) : reduce_variance_1d(in, axis), in); };
  if (nullptr == training) {
    get_mean = batch_mean;
    get_var = batch_var;
  } else {
    const auto dtype = (_GENERATED_DTYPE)input->get_meta().type_code();
    if (nullptr == momentum) momentum = eteq::make_constant_like(0.99, input);
    if (!moving_mean_init) moving_mean_init = init::zeros(dtype);
    if (!moving_var_init) moving_var_init = init::ones(dtype);
    VarptrT moving_mean = moving_mean_init(input->shape(), "moving_mean");
    VarptrT moving_var = moving_var_init(input->shape(), "moving_var");
    if (false == training->shape().compatible_after(input->shape(), 0)) training = extend_like(training, input);
    auto moving = [training, momentum](const VarptrT& state, std::function<ETensor(const ETensor&)> batch_stat) {
      return [=](const ETensor& in) {
        auto stat = batch_stat(in);
        auto blended = add(mul(ETensor(state), momentum), mul(stat, sub(1., momentum)));
        return if_then_else(training, stat, assign(state, blended));
      };
    };
    get_mean = moving(moving_mean, batch_mean);
    get_var = moving(moving_var, batch_var);
  }
  return nn::batch_normalization(input, offset, scale, eps, get_mean, get_var);
}

ETensor bind(layr::UnaryF unary, const Shape& inshape, _GENERATED_DTYPE dtype) {  // layer.yml:14-30
  ETensor input = eteq::make_variable_scalar(0, inshape, layr::input_label, dtype);
  auto output = unary(input);
  return layr::make_layer(identity(output), layr::bind_name, input);
}

ETensor link(ETensorsT layers, ETensor input) {  // layer.yml:31-67
  if (layers.empty()) global::fatal("cannot link without layers");
  ETensor output = input;
  if (nullptr == input) {
    output = layers.front();
    input = layr::get_input(output);
    layers = ETensorsT(layers.begin() + 1, layers.end());
  }
  for (auto& layer : layers) {
    if (layr::get_input(layer).get() == output.get()) output = layer;
    else output = layr::connect(layer, output);
  }
  return layr::make_layer(identity(output), layr::link_name, input);
}

ETensor dense(const ETensor& input, const ETensor& kernel, const ETensor& bias, eigen::PairVecT<RankT> dims) {  // layer.yml:636-656
  auto output = nn::fully_connect({input}, {kernel}, bias, dims);
  return layr::make_layer(identity(output), layr::dense_name, input);
}

ETensor dense(const ETensor& input, const DimsT& hidden_dims, layr::InitF kernel_init, layr::InitF bias_init, bool with_bias,
              const eigen::PairVecT<RankT>& dims) {  // layer.yml:89-129
  auto dtype = (_GENERATED_DTYPE)input->get_meta().type_code();
  if (!kernel_init) kernel_init = init::glorot_uniform(1, dtype);
  VarptrT kernel = kernel_init(layr::gen_rshape(hidden_dims, input->shape(), dims), layr::weight_label);
  VarptrT bias;
  if (with_bias) {
    if (!bias_init) bias_init = init::zeros(dtype);
    bias = bias_init(Shape(hidden_dims), layr::bias_label);
  }
  return dense(input, kernel, bias, dims);
}

ETensor dense(const Shape& inshape, const DimsT& hidden_dims, layr::InitF kernel_init, layr::InitF bias_init, bool with_bias,
              const eigen::PairVecT<RankT>& dims, _GENERATED_DTYPE dtype) {  // layer.yml:68-88
  ETensor input = eteq::make_variable_scalar(0, inshape, layr::input_label, dtype);
  return dense(input, hidden_dims, kernel_init, bias_init, with_bias, dims);
}

ETensor conv2d(const ETensor& input, const ETensor& kernel, const ETensor& bias, const std::pair<DimPairsT, DimPairsT>& zero_padding) {  // layer.yml:657-677
  auto output = nn::conv2d(input, kernel, bias, zero_padding);
  return layr::make_layer(identity(output), layr::conv_name, input);
}

ETensor conv2d(const DimPairsT& kernel_hw, DimT in_ncol, DimT out_ncol, layr::InitF kernel_init, layr::InitF bias_init,
               const std::pair<DimPairsT, DimPairsT>& zero_padding, bool with_bias, _GENERATED_DTYPE dtype) {  // layer.yml:130-160,215-253
  ETensor input = eteq::make_variable_scalar(0, Shape({in_ncol, kernel_hw.second, kernel_hw.first, 1}), layr::input_label, dtype);
  if (!kernel_init) kernel_init = init::glorot_uniform(1, dtype);
  VarptrT kernel = kernel_init(Shape({out_ncol, input->shape().at(0), kernel_hw.second, kernel_hw.first}), layr::weight_label);
  VarptrT bias;
  if (with_bias) {
    if (!bias_init) bias_init = init::zeros(dtype);
    bias = bias_init(Shape({out_ncol}), layr::bias_label);
  }
  return conv2d(input, kernel, bias, zero_padding);
}

ETensor conv2d(const ETensor& input, DimT out_ncol, const DimPairsT& kernel_hw, layr::InitF kernel_init, layr::InitF bias_init,
               const std::pair<DimPairsT, DimPairsT>& zero_padding, bool with_bias) {  // layer.yml:209-251: image must be [in, iwidth, iheight, ...]
  auto dtype = (_GENERATED_DTYPE)input->get_meta().type_code();
  if (!kernel_init) kernel_init = init::glorot_uniform(1, dtype);
  VarptrT kernel = kernel_init(Shape({out_ncol, input->shape().at(0), kernel_hw.second, kernel_hw.first}), layr::weight_label);
  VarptrT bias;
  if (with_bias) {
    if (!bias_init) bias_init = init::zeros(dtype);
    bias = bias_init(Shape({out_ncol}), layr::bias_label);
  }
  return conv2d(input, kernel, bias, zero_padding);
}

ETensor conv2d(const ETensor& input, DimT out_ncol, const DimPairsT& kernel_hw, layr::InitF kernel_init, layr::InitF bias_init,
               const std::string& padding, bool with_bias) {  // layer.yml:159-208
  std::string zpadding;
  std::transform(padding.begin(), padding.end(), std::back_inserter(zpadding), [](unsigned char c) { return std::tolower(c); });
  std::pair<DimPairsT, DimPairsT> zero_padding{{0, 0}, {0, 0}};
  if ("same" == zpadding) {
    DimT xpad = kernel_hw.second / 2, ypad = kernel_hw.first / 2;
    zero_padding = {{xpad, xpad}, {ypad, ypad}};
  } else if ("valid" != zpadding) {
    global::fatalf("unsupported padding type %s", padding.c_str());
  }
  return conv2d(input, out_ncol, kernel_hw, kernel_init, bias_init, zero_padding, with_bias);
}

static void check_seq_dim(RankT seq_dim) {
  if (seq_dim == 0) global::fatal("spliting input across 0th dimension... dense connection will not match");
}

ETensor rnn(const ETensor& input, const ETensor& init_state, const ETensor& cell, const layr::UnaryF& activation, RankT seq_dim) {  // layer.yml:678-715
  DimT nseq = input->shape().at(seq_dim);
  check_seq_dim(seq_dim);
  ETensor state = init_state;
  ETensorsT states;
  for (DimT i = 0; i < nseq; ++i) {
    ETensor inslice = slice(input, i, 1, seq_dim);
    state = activation(layr::connect(cell, concat(inslice, state, 0)));
    states.push_back(state);
  }
  auto output = concat(states, seq_dim);
  return layr::make_layer(identity(output), layr::rnn_name, input);
}

ETensor rnn(DimT indim, DimT hidden_dim, const layr::UnaryF& activation, DimT nseq, layr::InitF kernel_init, layr::InitF bias_init,
            RankT seq_dim, bool with_bias, _GENERATED_DTYPE dtype) {  // layer.yml:254-297
  DimsT inslist(rank_cap, 1);
  inslist[0] = indim;
  inslist[seq_dim] = nseq;
  ETensor input = eteq::make_variable_scalar(0, Shape(inslist), layr::input_label, dtype);
  auto cell = dense(Shape({(DimT)(hidden_dim + indim)}), {hidden_dim}, kernel_init, bias_init, with_bias, {{0, 1}}, dtype);
  auto init_state = eteq::make_variable_scalar(0, Shape({hidden_dim}), "init_state", dtype);
  ETensor state = extend_like(init_state, slice(input, 0, 1, seq_dim));
  return rnn(input, state, cell, activation, seq_dim);
}

ETensor lstm(const ETensor& input, const ETensor& init_state, const ETensor& init_hidden, const ETensor& ggate, const ETensor& forgate,
             const ETensor& ingate, const ETensor& outgate, RankT seq_dim) {  // layer.yml:716-768
  DimT nseq = input->shape().at(seq_dim);
  check_seq_dim(seq_dim);
  ETensor state = init_state, hidden = init_hidden;
  ETensorsT states;
  for (DimT i = 0; i < nseq; ++i) {
    ETensor inslice = slice(input, i, 1, seq_dim);
    ETensor xc = concat(inslice, hidden, 0);
    auto gate = tanh(layr::connect(ggate, xc));
    auto in = sigmoid(layr::connect(ingate, xc));
    auto forget = sigmoid(layr::connect(forgate, xc));
    auto output = sigmoid(layr::connect(outgate, xc));
    state = add(mul(gate, in), mul(state, forget));
    hidden = mul(state, output);
    states.push_back(hidden);
  }
  auto output = concat(states, seq_dim);
  return layr::make_layer(identity(output), layr::lstm_name, input);
}

ETensor lstm(const Shape& inshape, DimT hidden_dim, DimT nseq, layr::InitF kernel_init, layr::InitF bias_init, RankT seq_dim, bool with_bias,
             _GENERATED_DTYPE dtype) {  // layer.yml:298-345
  DimsT inslist(inshape.begin(), inshape.end());
  inslist[seq_dim] = nseq;
  ETensor input = eteq::make_variable_scalar(0, Shape(inslist), layr::input_label, dtype);
  DimsT inputlist(inshape.begin(), inshape.end()), statelist(inshape.begin(), inshape.end());
  inputlist[0] += hidden_dim;
  statelist[0] = hidden_dim;
  inputlist[seq_dim] = statelist[seq_dim] = 1;
  Shape inputshape(inputlist), stateshape(statelist);
  DimsT hid_dims = {hidden_dim};
  auto ggate = dense(inputshape, hid_dims, kernel_init, bias_init, with_bias, {{0, 1}}, dtype);
  auto forgate = dense(inputshape, hid_dims, kernel_init, bias_init, with_bias, {{0, 1}}, dtype);
  auto ingate = dense(inputshape, hid_dims, kernel_init, bias_init, with_bias, {{0, 1}}, dtype);
  auto outgate = dense(inputshape, hid_dims, kernel_init, bias_init, with_bias, {{0, 1}}, dtype);
  auto state = eteq::make_constant_scalar(0, stateshape, dtype);
  auto hidden = eteq::make_constant_scalar(0, stateshape, dtype);
  return lstm(input, state, hidden, ggate, forgate, ingate, outgate, seq_dim);
}

ETensor gru(const ETensor& input, const ETensor& init_state, const ETensor& ugate, const ETensor& rgate, const ETensor& hgate, RankT seq_dim) {  // layer.yml:769-813
  DimT nseq = input->shape().at(seq_dim);
  check_seq_dim(seq_dim);
  ETensor state = init_state;
  ETensorsT states;
  for (DimT i = 0; i < nseq; ++i) {
    ETensor inslice = slice(input, i, 1, seq_dim);
    ETensor xc = concat(inslice, state, 0);
    auto update = sigmoid(layr::connect(ugate, xc));
    auto reset = sigmoid(layr::connect(rgate, xc));
    auto hidden = tanh(layr::connect(hgate, concat(inslice, mul(reset, state), 0)));
    state = add(mul(update, state), mul(sub(1.0, update), hidden));
    states.push_back(state);
  }
  auto output = concat(states, seq_dim);
  return layr::make_layer(identity(output), layr::gru_name, input);
}

ETensor gru(const Shape& inshape, DimT hidden_dim, DimT nseq, layr::InitF kernel_init, layr::InitF bias_init, RankT seq_dim, bool with_bias,
            _GENERATED_DTYPE dtype) {  // layer.yml:346-390
  DimsT inslist(inshape.begin(), inshape.end());
  inslist[seq_dim] = nseq;
  ETensor input = eteq::make_variable_scalar(0, Shape(inslist), layr::input_label, dtype);
  DimsT inputlist(inshape.begin(), inshape.end()), statelist(inshape.begin(), inshape.end());
  inputlist[0] += hidden_dim;
  statelist[0] = hidden_dim;
  inputlist[seq_dim] = statelist[seq_dim] = 1;
  Shape inputshape(inputlist), stateshape(statelist);
  DimsT hid_dims = {hidden_dim};
  auto ugate = dense(inputshape, hid_dims, kernel_init, bias_init, with_bias, {{0, 1}}, dtype);
  auto rgate = dense(inputshape, hid_dims, kernel_init, bias_init, with_bias, {{0, 1}}, dtype);
  auto hgate = dense(inputshape, hid_dims, kernel_init, bias_init, with_bias, {{0, 1}}, dtype);
  auto state = eteq::make_constant_scalar(0, stateshape, dtype);
  return gru(input, state, ugate, rgate, hgate, seq_dim);
}

layr::RBMLayer rbm(DimT nvisible, DimT nhidden, layr::InitF kernel_init, layr::InitF bias_init, bool with_bias, _GENERATED_DTYPE dtype) {  // layer.yml:391-430
  if (!kernel_init) kernel_init = init::glorot_uniform(1, dtype);
  ETensor fwdinput = eteq::make_variable_scalar(0, Shape({nvisible}), layr::input_label, dtype);
  ETensor bwdinput = eteq::make_variable_scalar(0, Shape({nhidden}), layr::input_label, dtype);
  VarptrT kernel = kernel_init(Shape({nhidden, nvisible}), layr::weight_label);
  VarptrT hbias, vbias;
  if (with_bias) {
    if (!bias_init) bias_init = init::zeros(dtype);
    hbias = bias_init(Shape({nhidden}), "h" + layr::bias_label);
    vbias = bias_init(Shape({nvisible}), "v" + layr::bias_label);
  }
  return layr::RBMLayer{dense(fwdinput, kernel, hbias, {{0, 1}}), dense(bwdinput, transpose(kernel), vbias, {{0, 1}})};
}

}  // namespace layer

// ------------------------------------------------------------------ loss.yml
namespace loss {
ETensor sqr_diff(const ETensor& target, const ETensor& input) { return square(sub(target, input)); }
ETensor mean_squared(const ETensor& target, const ETensor& input, RankT axis) {  // loss.yml:21-39
  auto sd = square(sub(target, input));
  if (axis >= rank_cap) return reduce_mean(sd);
  return reduce_mean_1d(sd, axis);
}
ETensor cross_entropy(const ETensor& target, const ETensor& input, float eps) {  // loss.yml:40-59
  auto in = add(input, (double)eps);
  auto not_in = sub(1.0, in);
  auto not_targ = sub(1.0, target);
  return neg(add(mul(target, log(in)), mul(not_targ, log(not_in))));
}
}  // namespace loss

ETensorsT derive(const ETensor& root, const ETensorsT& targets) {
  return dp::wrap_gradients(eteq::derive(root, targets));
}

// ------------------------------------------------------------------ approx.yml
namespace approx {

static ETensorsT ders_of(const ETensor& error, const eteq::VarptrsT& variables) {
  return tenncor::derive(error, ETensorsT(variables.begin(), variables.end()));
}

static _GENERATED_DTYPE dtype_of(const ETensor& t) { return (_GENERATED_DTYPE)t->get_meta().type_code(); }

// The reference's optimizers are templates on the element type T and their hyper-parameters are of type T: a float model holds
// decay = 0.999f, and `T nodecay = 1. - decay` is 0.000999987, not 0.001 (tenncor/test/test_approx.cpp:388-482 prints it). Scalar
// arithmetic on hyper-parameters therefore happens here in the element type too.
static double in_elem(double v, _GENERATED_DTYPE dt) { return dt == FLOAT ? (double)(float)v : v; }
static double one_minus(double v, _GENERATED_DTYPE dt) { return in_elem(1. - in_elem(v, dt), dt); }

layr::VarErrsT sgd(const ETensor& error, const eteq::VarptrsT& variables, double learning_rate, layr::UnaryF apply) {  // approx.yml:12-55
  layr::VarErrsT out;
  auto ders = ders_of(error, variables);
  for (size_t i = 0, n = variables.size(); i < n; ++i) {
    auto der = ders[i];
    if (apply) der = apply(der);
    out.push_back({variables[i], assign_sub(variables[i], mul(der, learning_rate))});
  }
  return out;
}

layr::VarErrsT adagrad(const ETensor& error, const eteq::VarptrsT& variables, double learning_rate, double epsilon, layr::UnaryF apply) {  // approx.yml:56-95
  layr::VarErrsT out;
  auto ders = ders_of(error, variables);
  for (size_t i = 0, n = variables.size(); i < n; ++i) {
    auto der = ders[i];
    if (apply) der = apply(der);
    VarptrT momentum = eteq::make_variable_scalar(1, der->shape(), "momentum", dtype_of(der));
    auto update = assign_add(momentum, square(der));
    // assign momentums before leaves
    out.push_back({variables[i], assign_sub(variables[i], div(mul(der, learning_rate), add(sqrt(update), epsilon)))});
  }
  return out;
}

layr::VarErrsT adam(const ETensor& error, const eteq::VarptrsT& variables, double step_rate, double decay1, double decay2, double epsilon) {  // approx.yml:96-170
  layr::VarErrsT out;
  auto ders = ders_of(error, variables);
  for (size_t i = 0, n = variables.size(); i < n; ++i) {
    auto der = ders[i];
    auto dt = dtype_of(der);
    const double nodecay1 = one_minus(decay1, dt), nodecay2 = one_minus(decay2, dt);  // T nodecay = 1. - decay
    auto m = eteq::make_variable_scalar(0, der->shape(), "moment1", dt);
    auto v = eteq::make_variable_scalar(0, der->shape(), "moment2", dt);
    auto t = eteq::make_variable_scalar(0, der->shape(), "t", dt);
    auto one_t = make_constant_like(1, der);
    auto next_m = assign(m, add(mul(decay1, ETensor(m)), mul(nodecay1, der)));
    auto next_v = assign(v, add(mul(decay2, ETensor(v)), mul(nodecay2, square(der))));
    auto incr = assign_add(t, one_t);
    auto m_corr = div(next_m, sub(one_t, pow(decay1, incr)));
    auto v_corr = div(next_v, sub(one_t, pow(decay2, incr)));
    auto delta = mul(step_rate, div(m_corr, add(sqrt(v_corr), epsilon)));
    out.push_back({variables[i], assign_sub(variables[i], delta)});
  }
  return out;
}

layr::VarErrsT adadelta(const ETensor& error, const eteq::VarptrsT& variables, double step_rate, double decay, double offset, double epsilon,
                        layr::UnaryF apply) {  // approx.yml:171-245
  layr::VarErrsT out;
  auto ders = ders_of(error, variables);
  for (size_t i = 0, n = variables.size(); i < n; ++i) {
    auto der = ders[i];
    if (apply) der = apply(der);
    auto dt = dtype_of(der);
    const double nodecay = one_minus(decay, dt);  // T nodecay = 1. - decay (approx.yml:233)
    VarptrT msg = eteq::make_variable_scalar(0, der->shape(), "ex_sqr_grad", dt);
    VarptrT msd = eteq::make_variable_scalar(0, der->shape(), "ex_sqr_delx", dt);
    auto msg_next = assign(msg, add(mul(decay, ETensor(msg)), mul(nodecay, square(der))));
    auto delta = mul(mul(step_rate, div(sqrt(add(ETensor(msd), offset)), add(sqrt(add(msg_next, offset)), epsilon))), der);
    auto msd_next = assign(msd, add(mul(decay, ETensor(msd)), mul(nodecay, square(delta))));
    out.push_back({variables[i], assign_sub(variables[i], identity(delta, {msd_next}))});
  }
  return out;
}

layr::VarErrsT rms_momentum(const ETensor& error, const eteq::VarptrsT& variables, double learning_rate, double discount_factor,
                            double epsilon, layr::UnaryF apply) {  // approx.yml:246-322
  layr::VarErrsT out;
  auto ders = ders_of(error, variables);
  for (size_t i = 0, n = variables.size(); i < n; ++i) {
    auto der = ders[i];
    if (apply) der = apply(der);
    VarptrT momentum = eteq::make_variable_scalar(1, der->shape(), "momentum", dtype_of(der));
    auto update = assign(momentum, add(mul(discount_factor, ETensor(momentum)), mul(one_minus(discount_factor, dtype_of(der)), square(der))));
    // assign momentums before leaves
    out.push_back({variables[i], assign_sub(variables[i], div(mul(der, learning_rate), add(sqrt(update), epsilon)))});
  }
  return out;
}

}  // namespace approx

}  // namespace tenncor

// ======================================================================== trainer
namespace trainer {

using namespace tenncor;

layr::ETensor apply_update(const layr::ETensorsT& models, layr::ApproxF update, layr::ErrorF err_func) {
  auto error = err_func(models);
  eteq::VarptrsT vars;
  for (auto& model : models) {
    auto temp_vars = layr::get_storage(model);
    vars.insert(vars.end(), temp_vars.begin(), temp_vars.end());
  }
  auto updates = update(error, vars);
  teq::OwnMapT umap;
  layr::ETensorsT deps;
  deps.reserve(updates.size());
  for (auto& u : updates) {
    umap.emplace(u.first.get(), u.second);
    deps.push_back(u.second);
  }
  // depend on assigns for variables not trailed in error
  return identity(layr::trail(error, umap), deps);
}

layr::ETensor sample_v2h(const layr::RBMLayer& model, layr::ETensor vis) { return random::rand_binom_one(sigmoid(model.connect(vis))); }
layr::ETensor sample_h2v(const layr::RBMLayer& model, layr::ETensor hid) { return random::rand_binom_one(sigmoid(model.backward_connect(hid))); }
layr::ETensor gibbs_hvh(const layr::RBMLayer& model, layr::ETensor hid) { return sample_v2h(model, sample_h2v(model, hid)); }

layr::VarErrsT bbernoulli_approx(const layr::VarErrsT& assocs, double learning_rate, double discount_factor) {  // rbm.hpp:42-63
  layr::VarErrsT assigns;
  for (const auto& verrs : assocs) {
    auto err = verrs.second;
    auto slist = teq::narrow_shape(err->shape());
    teq::DimT shape_factor = slist.empty() ? 1 : slist.back();
    const auto dt = (egen::_GENERATED_DTYPE)err->get_meta().type_code();
    auto momentum = eteq::make_variable_scalar(0, err->shape(), "momentum", dt);
    // (learning_rate * (1 - discount_factor) / shape_factor) with T-typed hyper-parameters: every step rounds to T (rbm.hpp:55-56)
    using tenncor::approx::in_elem;
    using tenncor::approx::one_minus;
    const double step = in_elem(in_elem(in_elem(learning_rate, dt) * one_minus(discount_factor, dt), dt) / in_elem((double)shape_factor, dt), dt);
    auto momentum_next = add(mul(discount_factor, layr::ETensor(momentum)), mul(step, err));
    assigns.push_back({verrs.first, assign_add(verrs.first, assign(momentum, momentum_next))});
  }
  return assigns;
}

layr::VarErrsT cd_grad_approx(CDChainIO& io, const layr::RBMLayer& model, size_t cdk, eteq::VarptrT persistent) {  // rbm.hpp:83-146
  if (nullptr == io.visible_) global::fatal("cannot call cd_grad_approx with null visible");
  if (nullptr == io.hidden_) io.hidden_ = sample_v2h(model, io.visible_);
  layr::ETensor chain_it = nullptr == persistent ? io.hidden_ : layr::ETensor(persistent);
  for (size_t i = 0; i + 1 < cdk; ++i) chain_it = gibbs_hvh(model, chain_it);
  io.visible_mean_ = sigmoid(model.backward_connect(chain_it));
  io.hidden_mean_ = sigmoid(model.connect(io.visible_mean_));

  std::map<std::string, eteq::VarptrT> vars;
  for (auto& var : layr::get_storage(model.fwd_)) vars.emplace(var->to_string(), var);
  for (auto& var : layr::get_storage(model.bwd_)) vars.emplace(var->to_string(), var);

  auto grad_w = sub(matmul(transpose(io.visible_), io.hidden_), matmul(transpose(io.visible_mean_), io.hidden_mean_));
  layr::VarErrsT varerrs = {{vars.at(layr::weight_label), grad_w}};
  const std::string hid_key = "h" + layr::bias_label, vis_key = "v" + layr::bias_label;
  if (vars.count(hid_key)) varerrs.push_back({vars.at(hid_key), reduce_mean_1d(sub(io.hidden_, io.hidden_mean_), 1)});
  if (vars.count(vis_key)) varerrs.push_back({vars.at(vis_key), reduce_mean_1d(sub(io.visible_, io.visible_mean_), 1)});
  if (nullptr != persistent) varerrs.push_back({persistent, gibbs_hvh(model, chain_it)});
  return varerrs;
}

layr::ETensor rbm(const layr::RBMLayer& model, layr::ETensor visible, double learning_rate, double discount_factor, BErrorF err_func, size_t cdk) {  // rbm.hpp:148-165
  if (!err_func) err_func = [](const layr::ETensor& a, const layr::ETensor& b) { return loss::mean_squared(a, b); };
  CDChainIO io(visible);
  layr::VarErrsT varerrs = cd_grad_approx(io, model, cdk);
  // the exchanged "errors" of a data-parallel RBM are the CD statistics
  layr::ETensorsT errs;
  for (auto& ve : varerrs) errs.push_back(ve.second);
  errs = dp::wrap_gradients(errs);
  for (size_t i = 0; i < errs.size(); ++i) varerrs[i].second = errs[i];
  auto updates = bbernoulli_approx(varerrs, learning_rate, discount_factor);
  teq::OwnMapT umap;
  for (auto& u : updates) umap.emplace(u.first.get(), u.second);
  layr::ETensor error = err_func(io.visible_, io.visible_mean_);
  return layr::trail(error, umap);
}

// ---------------------------------------------------------------- DBN (tenncor/trainer/dbn.hpp)
static double scalar_of(const layr::ETensor& t) {
  const void* data = t->device().data();
  if (nullptr == data) global::fatalf("%s has no data", t->to_string().c_str());
  double out = 0;
  type_convert(&out, DOUBLE, data, (_GENERATED_DTYPE)t->get_meta().type_code(), 1);
  return out;
}

DBNTrainer::DBNTrainer(const std::vector<layr::RBMLayer>& rbms, layr::ETensor dense, RankT softmax_dim, DimT batch_size, double pretrain_lr,
                       double train_lr, size_t cdk, double l2_reg, double lr_scaling)
    : nlayers_(rbms.size()), batch_size_(batch_size) {
  if (rbms.empty()) global::fatal("cannot train a deep belief network without rbm layers");
  input_size_ = layr::get_input(rbms.front().fwd_)->shape().at(0);
  output_size_ = dense->shape().at(0);
  const auto dtype = (_GENERATED_DTYPE)dense->get_meta().type_code();
  trainx_ = eteq::make_variable_scalar(0, Shape({(DimT)input_size_, batch_size}), "trainx", dtype);
  trainy_ = eteq::make_variable_scalar(0, Shape({(DimT)output_size_, batch_size}), "trainy", dtype);

  // general rbm sampling: every layer feeds on a sample of the one below
  sample_pipes_.push_back(trainx_);
  for (size_t i = 0; i < nlayers_; ++i) sample_pipes_.push_back(sample_v2h(rbms[i], sample_pipes_[i]));
  // the samples are evaluated once and then read as data by many later evaluations (pretrain / finetune name them in `ignored`):
  // their buffers must not expire with their consumers' reads — the reference's cache_init() (dbn.hpp:138-144,188-191)
  for (size_t i = 1; i < sample_pipes_.size(); ++i)
    if (auto op = dynamic_cast<cuda::DevOp*>(&sample_pipes_[i]->device())) op->pin();

  // layer-wise rbm reconstruction
  for (size_t i = 0; i < nlayers_; ++i) {
    const layr::RBMLayer& layer = rbms[i];
    const layr::ETensor& rx = sample_pipes_[i];
    teq::TensSetT to_learn;
    for (auto& var : layr::get_storage(layer.fwd_)) to_learn.emplace(var.get());
    for (auto& var : layr::get_storage(layer.bwd_)) to_learn.emplace(var.get());
    CDChainIO io(rx, sample_pipes_[i + 1]);
    layr::VarErrsT varerrs = cd_grad_approx(io, layer, cdk);
    layr::ETensorsT assigns;
    for (auto& varerr : varerrs)  // weights and biases move by the learning rate; anything else (a persistent chain) is replaced
      assigns.push_back(to_learn.count(varerr.first.get()) ? assign_add(varerr.first, mul(pretrain_lr, varerr.second)) : assign(varerr.first, varerr.second));
    rupdates_.push_back(assigns);
    auto vhv = sigmoid(layer.backward_connect(sigmoid(layer.connect(rx))));
    rcosts_.push_back(neg(reduce_mean(reduce_sum_1d(add(mul(rx, log(vhv)), mul(sub(1., rx), log(sub(1., vhv)))), 0))));
  }

  // logistic layer on the top-level samples
  auto contents = layr::get_storage(dense);
  if (contents.size() < 2) global::fatal("the dbn's dense layer needs a weight and a bias");
  eteq::VarptrT w = contents[0], b = contents[1];
  auto final_out = softmax(layr::connect(dense, sample_pipes_.back()), softmax_dim, 1);
  auto diff = sub(layr::ETensor(trainy_), final_out);
  auto l2_regularized = sub(matmul(transpose(sample_pipes_.back()), diff), mul(l2_reg, layr::ETensor(w)));
  Shape wshape = w->shape(), bshape = b->shape();
  auto tlr = eteq::make_variable_scalar(train_lr, Shape(), "learning_rate", dtype);
  auto dw = mul(extend(layr::ETensor(tlr), 0, DimsT(wshape.begin(), wshape.end())), l2_regularized);
  auto db = mul(extend(layr::ETensor(tlr), 0, DimsT(bshape.begin(), bshape.end())), reduce_mean_1d(diff, 1));
  auto dtrain_lr = mul(layr::ETensor(tlr), lr_scaling);
  tupdate_ = assign(tlr, identity(dtrain_lr, {assign_add(w, dw), assign_add(b, db)}));
  tcost_ = neg(reduce_mean(reduce_sum_1d(add(mul(layr::ETensor(trainy_), log(final_out)), mul(sub(1., layr::ETensor(trainy_)), log(sub(1., final_out)))), 0)));
}

void DBNTrainer::pretrain(const void* train_in, _GENERATED_DTYPE dtype, size_t nepochs, std::function<void(size_t, size_t)> logger) {
  trainx_->assign(train_in, dtype, trainx_->shape());
  for (size_t i = 0; i < nlayers_; ++i) {
    // the layer's input sample is frozen (ignored = read as data) while its RBM learns to reconstruct it
    teq::TensSetT ignore = {sample_pipes_[i].get()};
    for (size_t epoch = 0; epoch < nepochs; ++epoch) {
      eteq::run(rupdates_[i], ignore);
      if (logger) logger(epoch, i);
    }
    if (i + 1 < nlayers_) eteq::run({sample_pipes_[i + 1]}, ignore);
  }
}

void DBNTrainer::finetune(const void* train_in, const void* train_out, _GENERATED_DTYPE dtype, size_t nepochs, std::function<void(size_t)> logger) {
  trainx_->assign(train_in, dtype, trainx_->shape());
  trainy_->assign(train_out, dtype, trainy_->shape());
  layr::ETensor top = sample_pipes_.back();
  if (nullptr == sample_pipes_[nlayers_ - 1]->device().device_data() && nlayers_ > 1)
    global::fatal("finetune needs the samples pretrain leaves behind: call pretrain first");
  eteq::run({top}, nlayers_ > 1 ? teq::TensSetT{sample_pipes_[nlayers_ - 1].get()} : teq::TensSetT{});
  for (size_t epoch = 0; epoch < nepochs; ++epoch) {
    eteq::run({tupdate_}, {top.get()});  // train the logistic layer on the frozen top-level sample
    if (logger) logger(epoch);
  }
}

double DBNTrainer::reconstruction_cost(size_t layer) {
  if (layer >= rcosts_.size()) global::fatalf("layer %d out of range (%d rbm layers)", (int)layer, (int)rcosts_.size());
  eteq::run({rcosts_[layer]});
  return scalar_of(rcosts_[layer]);
}

double DBNTrainer::training_cost() {
  eteq::run({tcost_});
  return scalar_of(tcost_);
}

}  // namespace trainer
