"""Draw an equation graph as an ascii tree (reference: dbg/print/tree.hpp PrettyTree, teq.hpp PrettyEquation, src/teq.cpp
ten_stream; python module dbg/python/print.cpp)."""
import builtins

INDENT = " "  # default_indent (the reference's tests render with '_')


def _text(t, showshape, showtype, showvers):
    out = (t.usage() + ":" if t.is_leaf() else "") + t.label()
    if showtype:
        out += "<" + t.type_label() + ">"
    if showshape:
        out += "[" + "\\".join(str(d) for d in t.teq_shape()) + "]"
    if showvers:
        out += ":version=%d" % t.get_version()
    return out


def graph_to_str(root, showshape=False, showtype=False, showvers=False, indent=INDENT):
    """Return graph of root tensor as string"""
    lines = []

    def rec(t, prefix):
        lines.append("(" + _text(t, showshape, showtype, showvers) + ")\n")
        kids = [] if t.is_leaf() else t.args()
        branch = prefix + indent + "`--"
        for i, kid in enumerate(kids):
            lines.append(branch)
            last = i == len(kids) - 1
            rec(kid, prefix + (indent * 4 if last else indent + "|" + indent * 2))

    rec(root, "")
    return "".join(lines)


def print_graph(root, showshape=False):
    """Print graph of root tensor to stdout"""
    builtins.print(graph_to_str(root, showshape), end="")
