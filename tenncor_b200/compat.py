"""Run scripts written for the reference's python modules unchanged: `install()` registers this package under the names those
scripts import (`import tenncor as tc`, `import extenncor.dqn_trainer as etc`, `import dbg.compare as cmp`, `from dbg.print import
graph_to_str`). Opt-in and process-local: nothing is aliased unless a caller asks for it.

    python -c "import tenncor_b200.compat as c; c.install(); import runpy; runpy.run_path('demo/gd_demo.py', run_name='__main__')"
"""
import sys


def install():
    import tenncor_b200
    from tenncor_b200 import dbg, extenncor
    aliases = {
        "tenncor": tenncor_b200,
        "extenncor": extenncor, "extenncor.dqn_trainer": extenncor.dqn_trainer, "extenncor.trainer_cache": extenncor.trainer_cache,
        "dbg": dbg, "dbg.compare": dbg.compare, "dbg.print": dbg.print,
    }
    for name, module in aliases.items():
        sys.modules.setdefault(name, module)
    return sorted(aliases)
