"""The reference's python debugging helpers over TEQ graphs (dbg/python/print.cpp, compare.cpp): `dbg.print`, `dbg.compare`."""
from . import compare, print  # noqa: F401,A004
