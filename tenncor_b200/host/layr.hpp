// layr.hpp — layer bookkeeping on top of eteq: a layer is an IDENTITY root carrying a
// "layer" LayerObj attribute that names its input; connect / trail re-instantiate the
// sub-graph on a new input, sharing the leaf variables.
// Mirrors tenncor/layr/layer.hpp:17-193 and tenncor/layr/src/layer.cpp:10-103.
#ifndef TCR_HOST_LAYR_HPP
#define TCR_HOST_LAYR_HPP

#include "eteq.hpp"

namespace layr {

using ETensor = teq::TensptrT;
using ETensorsT = teq::TensptrsT;
using UnaryF = std::function<ETensor(const ETensor&)>;
using InitF = std::function<eteq::VarptrT(teq::Shape, std::string)>;
using VarErrsT = std::vector<std::pair<eteq::VarptrT, ETensor>>;
using ApproxF = std::function<VarErrsT(const ETensor&, const eteq::VarptrsT&)>;
using ErrorF = std::function<ETensor(const ETensorsT&)>;

const std::string weight_label = "weight";
const std::string bias_label = "bias";
const std::string input_label = "input";
const std::string bind_name = "_UNARY_BIND";
const std::string link_name = "_LINK";
const std::string dense_name = "_DENSE_LAYER";
const std::string conv_name = "_CONV_LAYER";
const std::string rnn_name = "_RNN_LAYER";
const std::string lstm_name = "_LSTM_LAYER";
const std::string gru_name = "_GRU_LAYER";

/// shape of the right operand of a contraction given the un-contracted dims wanted
teq::Shape gen_rshape(teq::DimsT runcoms, teq::Shape left, eigen::PairVecT<teq::RankT> lrdims);

ETensor make_layer(ETensor root, const std::string& layername, ETensor input);
ETensor get_input(const ETensor& root);
/// copy everything on a path from inputs.first to root, replacing inputs.first by inputs.second
ETensor trail(const ETensor& root, const teq::OwnMapT& inputs);
ETensor connect(const ETensor& root, const ETensor& input);
ETensor deep_clone(const ETensor& root);
/// mutable leaves of a layer in traversal order, excluding its input
eteq::VarptrsT get_storage(const ETensor& root);

struct RBMLayer final {
  RBMLayer deep_clone() const { return RBMLayer{layr::deep_clone(fwd_), layr::deep_clone(bwd_)}; }
  ETensor connect(const ETensor& input) const { return layr::connect(fwd_, input); }
  ETensor backward_connect(const ETensor& hidden) const { return layr::connect(bwd_, hidden); }
  ETensor fwd_;
  ETensor bwd_;
};

}  // namespace layr

#endif  // TCR_HOST_LAYR_HPP
