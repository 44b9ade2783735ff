"""Static checks of bench.py that need no GPU.

A multi-GPU bench step contains collectives (VTC all-gather, gradient averaging, peer-memory barriers): any code that
runs extra steps on a subset of the ranks hangs the job without an error message. This happened once (an energy
measurement loop under `if rank == 0`), so the property is pinned here: inside run_ours no call of a step function may
sit under a condition on `rank`."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEP_CALLS = {"step", "e2e_step", "e2e_u8_step", "step_opt"}


TAINTED = {"rank"}


def _taint(fn):
    """Names whose VALUE depends on the rank (e.g. `meter = Meter() if rank == 0 else None`): a condition on them is a
    condition on the rank. Fixpoint over the assignments of the function."""
    names = {"rank"}
    changed = True
    while changed:
        changed = False
        for node in ast.walk(fn):
            if isinstance(node, ast.Assign) and any(isinstance(x, ast.Name) and x.id in names for x in ast.walk(node.value)):
                for t in node.targets:
                    for x in ast.walk(t):
                        if isinstance(x, ast.Name) and x.id not in names:
                            names.add(x.id)
                            changed = True
    return names


def _mentions_rank(node):
    return any(isinstance(x, ast.Name) and x.id in TAINTED for x in ast.walk(node))


def _step_calls(node):
    out = []
    for x in ast.walk(node):
        if isinstance(x, ast.Call):
            f = x.func
            if isinstance(f, ast.Name) and f.id in STEP_CALLS:
                out.append((f.id, x.lineno))
            # D.timed(fn, ...) runs fn as steps as well
            if isinstance(f, ast.Attribute) and f.attr == "timed":
                out.append(("D.timed", x.lineno))
    return out


def test_no_step_runs_on_a_subset_of_the_ranks():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    TAINTED.clear()
    TAINTED.update(_taint(fn))
    bad = []
    for node in ast.walk(fn):
        if isinstance(node, (ast.If, ast.While)) and _mentions_rank(node.test):
            for name, line in _step_calls(node):
                bad.append(f"{name} at bench.py:{line} under a condition on rank (line {node.lineno})")
        if isinstance(node, ast.IfExp) and _mentions_rank(node.test):
            for name, line in _step_calls(node):
                bad.append(f"{name} at bench.py:{line} in a conditional expression on rank")
    assert not bad, bad


def test_rank0_prints_after_every_rank_has_left_the_collectives():
    """The early `return` of the non-zero ranks must come after the last step / parity call of run_ours."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    TAINTED.clear()
    TAINTED.add("rank")
    ret_line = None
    for node in ast.walk(fn):
        if isinstance(node, ast.If) and _mentions_rank(node.test) and any(isinstance(b, ast.Return) for b in node.body):
            if "!=" in ast.get_source_segment(src, node.test):
                ret_line = node.lineno
    assert ret_line is not None
    later = [(n, l) for n, l in _step_calls(fn) if l > ret_line]
    assert not later, later
