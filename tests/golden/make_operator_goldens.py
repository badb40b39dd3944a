"""Writes tests/golden/operator_goldens.json.

The vectors below are the hard-coded inputs / expected outputs of the reference's own
operator unit tests, transcribed by hand from
/root/reference/internal/eigen/test/test_operator.cpp (line numbers in `cite`). They are
facts about the reference's behaviour (the tests assert them against the Eigen back end),
kept here so that the oracle and the CUDA kernels can be pinned without the reference
tree, which does not exist on the GPU box. Run:  python tests/golden/make_operator_goldens.py
"""
import json
import os

F = "internal/eigen/test/test_operator.cpp"
cases = []


def case(name, lines, op, inputs, expect, out_shape, attrs=None, dtype="double", out_dtype=None):
    cases.append({"name": name, "cite": "%s:%s" % (F, lines), "op": op, "dtype": dtype,
                  "out_dtype": out_dtype or dtype,
                  "inputs": [{"shape": s, "data": d} for s, d in inputs],
                  "attrs": attrs or {}, "out_shape": out_shape, "expect": expect})


# reductions over rank 1 of [3,2] (test_reduce, :25-66; TESTs :96-117)
red_in = ([3, 2], [2, 3, 4, 5, 6, 7])
case("reduce_sum", "25-66,96-99", "REDUCE_SUM", [red_in], [7, 9, 11], [3], {"rank_set": [1]})
case("reduce_prod", "25-66,102-105", "REDUCE_PROD", [red_in], [10, 18, 28], [3], {"rank_set": [1]})
case("reduce_min", "25-66,108-111", "REDUCE_MIN", [red_in], [2, 3, 4], [3], {"rank_set": [1]})
case("reduce_max", "25-66,114-117", "REDUCE_MAX", [red_in], [5, 6, 7], [3], {"rank_set": [1]})
# ArgMax (:120-185)
case("argmax_dim1", "120-166", "ARGMAX", [([3, 2], [2, 8, 4, 5, 6, 7])], [1, 0, 1], [3], {"rank": 1})
case("argmax_flat", "131-184", "ARGMAX", [([3, 2], [2, 8, 4, 5, 9, 7])], [4], [1], {"rank": 8})
# Extend (:186-236)
case("extend", "186-236", "EXTEND", [([3, 1, 2], [2, 8, 4, 5, 6, 7])],
     [2, 8, 4, 2, 8, 4, 2, 8, 4, 2, 8, 4, 5, 6, 7, 5, 6, 7, 5, 6, 7, 5, 6, 7], [3, 4, 2],
     {"dimensions": [1, 4]})
# Permute (:239-374)
perm_in = ([2, 2, 3], [2, 8, 4, 5, 6, 7, 1, 0, 9, 11, 10, 12])
perm_out = [2, 6, 9, 8, 7, 11, 4, 1, 10, 5, 0, 12]
case("permute_full_order", "245-288", "PERMUTE", [perm_in], perm_out, [3, 2, 2], {"ranks": [2, 0, 1, 3, 4, 5, 6, 7]})
case("permute_partial_order", "289-330", "PERMUTE", [perm_in], perm_out, [3, 2, 2], {"ranks": [2, 0, 1, 3, 4, 5]})
case("permute_transpose", "331-373", "PERMUTE", [([2, 3], [2, 8, 4, 5, 6, 7])], [2, 4, 6, 8, 5, 7], [3, 2],
     {"ranks": [1, 0, 2, 3, 4, 5, 6, 7]})
# Slice (:377-441)
case("slice_box", "386-421", "SLICE", [([3, 2], [2, 8, 4, 5, 6, 7])], [6, 7], [2, 1], {"dimension_pairs": [[1, 2], [1, 1]]})
case("slice_lastdim_view", "424-439", "SLICE", [([3, 2], [2, 8, 4, 5, 6, 7])], [5, 6, 7], [3, 1], {"dimension_pairs": [[0, 3], [1, 1]]})
# MultiConcat (:444-560)
case("concat_nary2", "453-500", "CONCAT", [([1, 4], [2, 8, 4, 5]), ([1, 4], [1, 0, 3, 9])],
     [2, 1, 8, 0, 4, 3, 5, 9], [2, 4], {"rank": 0})
case("concat_nary3", "503-559", "CONCAT", [([1, 4], [2, 8, 4, 5]), ([1, 4], [1, 0, 3, 9]), ([1, 4], [3, 7, 2, 11])],
     [2, 1, 3, 8, 0, 7, 4, 3, 2, 5, 9, 11], [3, 4], {"rank": 0})
# Pow / Add / Sub / Mul / Div (:564-1142)
case("pow_2d", "570-612", "POW", [([2, 3], [2, 8, 4, 5, 6, 7]), ([2, 3], [1, 0, 3, 3, 2, 4])], [2, 1, 64, 125, 36, 2401], [2, 3])
case("pow_3d", "615-656", "POW", [([2, 2, 2], [2, 8, 4, 5, 6, 7, 4, 2]), ([2, 2, 2], [1, 0, 3, 3, 2, 4, 2, 3])],
     [2, 1, 64, 125, 36, 2401, 16, 8], [2, 2, 2])
case("add_3d", "668-710", "ADD", [([2, 2, 2], [2, 8, 4, 5, 6, 7, 8, 11]), ([2, 2, 2], [1, 0, 3, 9, 10, 11, 6, 1.2])],
     [3, 8, 7, 14, 16, 18, 14, 12.2], [2, 2, 2])
case("add_2d", "712-758", "ADD", [([2, 3], [2, 8, 4, 5, 6, 7]), ([2, 3], [1, 0, 3, 9, 10, 11])], [3, 8, 7, 14, 16, 18], [2, 3])
case("add_nary3", "712-801", "ADD", [([2, 3], [2, 8, 4, 5, 6, 7]), ([2, 3], [1, 0, 3, 9, 10, 11]), ([2, 3], [4.2, 1, 7.1, 1, 2, 1.1])],
     [7.2, 9, 14.1, 15, 18, 19.1], [2, 3])
case("sub_2d", "812-854", "SUB", [([2, 3], [2, 8, 4, 5, 6, 7]), ([2, 3], [1, 0, 3, 9, 10, 11])], [1, 8, 1, -4, -4, -4], [2, 3])
case("sub_3d", "857-898", "SUB", [([2, 2, 2], [2, 8, 4, 5, 6, 7, 8, 11]), ([2, 2, 2], [1, 0, 3, 9, 10, 11, 6, 1.2])],
     [1, 8, 1, -4, -4, -4, 2, 9.8], [2, 2, 2])
case("mul_3d", "908-950", "MUL", [([2, 2, 2], [2, 8, 4, 5, 6, 7, 1.2, 3]), ([2, 2, 2], [1, 0, 3, 9, 10, 11, 2, 1.7])],
     [2, 0, 12, 45, 60, 77, 2.4, 5.1], [2, 2, 2])
case("mul_2d", "953-999", "MUL", [([2, 3], [2, 8, 4, 5, 6, 7]), ([2, 3], [1, 0, 3, 9, 10, 11])], [2, 0, 12, 45, 60, 77], [2, 3])
case("mul_nary3", "953-1043", "MUL", [([2, 3], [2, 8, 4, 5, 6, 7]), ([2, 3], [1, 0, 3, 9, 10, 11]), ([2, 3], [4, 1, 7, 1, 2, 1])],
     [8, 0, 84, 45, 120, 77], [2, 3])
case("div_2d", "1053-1095", "DIV", [([2, 3], [2, 8, 4, 5, 6, 7]), ([2, 3], [1, 0.5, 3, 9, 10, 11])],
     [2, 16, 4. / 3, 5. / 9, 0.6, 7. / 11], [2, 3])
case("div_3d", "1098-1140", "DIV", [([2, 2, 2], [2, 8, 4, 5, 6, 7, 1.2, 3]), ([2, 2, 2], [1, 0.5, 3, 9, 10, 11, 2, 1.7])],
     [2, 16, 4. / 3, 5. / 9, 0.6, 7. / 11, 0.6, 3 / 1.7], [2, 2, 2])
# comparisons, min, max (:1145-1722)
ca2, cb2 = ([2, 3], [2, 8, 4, 5, 6, 7]), ([2, 3], [1, 0.5, 4, 9, 6, 11])
ca3, cb3 = ([2, 2, 2], [2, 8, 4, 5, 6, 7, 3, 8]), ([2, 2, 2], [1, 0.5, 4, 9, 6, 11, 3, 3])
for nm, op, l2, l3, e2, e3 in [
    ("eq", "EQ", "1150-1192", "1195-1237", [0, 0, 1, 0, 1, 0], [0, 0, 1, 0, 1, 0, 1, 0]),
    ("neq", "NEQ", "1247-1288", "1291-1333", [1, 1, 0, 1, 0, 1], [1, 1, 0, 1, 0, 1, 0, 1]),
    ("lt", "LT", "1343-1385", "1388-1430", [0, 0, 0, 1, 0, 1], [0, 0, 0, 1, 0, 1, 0, 0]),
    ("gt", "GT", "1440-1481", "1484-1526", [1, 1, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0, 0, 1]),
    ("min", "MIN", "1536-1577", "1580-1623", [1, 0.5, 4, 5, 6, 7], [1, 0.5, 4, 5, 6, 7, 3, 3]),
    ("max", "MAX", "1633-1674", "1677-1720", [2, 8, 4, 9, 6, 11], [2, 8, 4, 9, 6, 11, 3, 8]),
]:
    case(nm + "_2d", l2, op, [ca2, cb2], e2, [2, 3])
    case(nm + "_3d", l3, op, [ca3, cb3], e3, [2, 2, 2])
# Select (:1854-1966)
case("select_2d", "1859-1910", "SELECT", [([2, 3], [0, 1, 0, 0, 1, 1]), ([2, 3], [2, 8, 9, 5, 8, 7]), ([2, 3], [1, 0.5, 4, 9, 6, 11])],
     [1, 8, 4, 9, 8, 7], [2, 3])
case("select_3d", "1913-1964", "SELECT", [([2, 2, 2], [0, 1, 0, 0, 1, 1, 0, 1]), ([2, 2, 2], [2, 8, 9, 5, 8, 7, 4, 8]), ([2, 2, 2], [1, 0.5, 4, 9, 6, 11, 3, 3])],
     [1, 8, 4, 9, 8, 7, 3, 8], [2, 2, 2])
# Contract (:1969-2078)
case("contract_2d", "1974-2020", "CONTRACT", [([4, 3], [2, 8, 9, 5, 8, 7, 1, 9, 4.2, 3, 2, 6]), ([2, 4], [1, 0.5, 4, 9, 6, 11, 3, 8])],
     [103, 212, 69, 150, 46.2, 99.1], [2, 3], {"rank_pairs": [[0, 1]]})
case("contract_3d", "2023-2076", "CONTRACT",
     [([4, 2, 3], [2, 8, 9, 5, 8, 7, 1, 9, 4.2, 3, 2, 6, 2, 8, 9, 5, 8, 7, 1, 9, 4.2, 3, 2, 6]),
      ([2, 4, 2], [1, 0.5, 4, 9, 6, 11, 3, 8, 1, 0.5, 4, 9, 6, 11, 3, 8])],
     [172, 362, 149.2, 311.1, 115.2, 249.1], [2, 3], {"rank_pairs": [[0, 1], [1, 2]]})
# Pad / Stride / Scatter / Reverse / Concat (:2081-2343)
case("pad", "2081-2129", "PAD", [([2, 3], [2, 8, 4, 5, 6, 7])], [0, 2, 8, 0, 0, 4, 5, 0, 0, 6, 7, 0], [4, 3], {"dimension_pairs": [[1, 1]]})
case("stride", "2134-2179", "STRIDE", [([2, 3], [2, 8, 4, 5, 6, 7])], [2, 8, 6, 7], [2, 2], {"dimensions": [1, 2]})
case("scatter", "2184-2231", "SCATTER", [([2, 2], [2, 8, 4, 5])], [2, 0, 8, 0, 0, 0, 4, 0, 5], [3, 3], {"dimensions": [2, 2], "shape": [3, 3]})
case("reverse", "2236-2281", "REVERSE", [([2, 3], [2, 8, 4, 5, 6, 7])], [6, 7, 4, 5, 2, 8], [2, 3], {"rank_set": [1]})
case("concat_binary", "2286-2341", "CONCAT", [([2, 3], [2, 8, 4, 5, 7, 6]), ([1, 3], [1, 0, 3])], [2, 8, 1, 4, 5, 0, 7, 6, 3], [3, 3], {"rank": 0})
# Convolution (:2538-2644): valid correlation along image rank 1
case("convolution", "2590-2640", "CONV", [([3, 3], [2, 8, 4, 5, 7, 6, 9, 1, 0]), ([2], [0.3, 0.6])],
     [2 * 0.3 + 5 * 0.6, 8 * 0.3 + 7 * 0.6, 4 * 0.3 + 6 * 0.6, 5 * 0.3 + 9 * 0.6, 7 * 0.3 + 1 * 0.6, 6 * 0.3 + 0 * 0.6], [3, 2], {"ranks": [1]})
# Assign* (:2647-2819)
asg_a, asg_b = ([2, 3], [2, 8, 4, 5, 6, 7]), ([2, 3], [1, 0, 3, 9, 10, 11])
case("assign", "2647-2678", "ASSIGN", [asg_a, asg_b], [1, 0, 3, 9, 10, 11], [2, 3])
case("assign_add", "2681-2712", "ASSIGN_ADD", [asg_a, asg_b], [3, 8, 7, 14, 16, 18], [2, 3])
case("assign_sub", "2716-2747", "ASSIGN_SUB", [asg_a, asg_b], [1, 8, 1, -4, -4, -4], [2, 3])
case("assign_mul", "2751-2782", "ASSIGN_MUL", [asg_a, asg_b], [2, 0, 12, 45, 60, 77], [2, 3])
case("assign_div", "2786-2817", "ASSIGN_DIV", [asg_a, ([2, 3], [1, 2, 3, 9, 10, 11])], [2, 4, 4. / 3, 5. / 9, 0.6, 7. / 11], [2, 3])
# Cast double -> int32 (:2821-2878)
case("cast_double_int32", "2821-2878", "CAST", [([2, 3], [2.1, 8.5, 4.3, 5.2, 6.1, 7.2])], [2, 8, 4, 5, 6, 7], [2, 3], out_dtype="int32")
# unary ops vs std:: functions on {-2, 8, -4, -5, 7, 6} in [2,3] and [2,1,3] (:2361-2535);
# expected values are computed by the test with std::abs / std::sin / ... at run time, so the
# fixture records the op + input and the checker applies the same libm function in double.
un_in = [-2, 8, -4, -5, 7, 6]
for nm in ["ABS", "NEG", "SIN", "COS", "TAN", "EXP", "SIGMOID", "TANH", "SQUARE", "CUBE"]:
    case("unary_%s_2d" % nm.lower(), "2361-2535", nm, [([2, 3], un_in)], None, [2, 3])
    case("unary_%s_3d" % nm.lower(), "2361-2535", nm, [([2, 1, 3], un_in)], None, [2, 1, 3])
# Log / Sqrt use positive inputs {3, 8, 2, 5, 7, 3}; Round uses {3.22, 8.51, 2.499, 5.2, 7.17, 3.79} (:2483-2502)
for nm, data in [("LOG", [3, 8, 2, 5, 7, 3]), ("SQRT", [3, 8, 2, 5, 7, 3]), ("ROUND", [3.22, 8.51, 2.499, 5.2, 7.17, 3.79])]:
    case("unary_%s_2d" % nm.lower(), "2483-2502", nm, [([2, 3], data)], None, [2, 3])

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "operator_goldens.json")
with open(out, "w") as f:
    json.dump({"source": "/root/reference/" + F, "cases": cases}, f, indent=1)
print("wrote %d cases to %s" % (len(cases), out))
