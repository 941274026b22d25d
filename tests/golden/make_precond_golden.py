"""Golden lines of the UNMODIFIED reference's preconditioned solves (row f-2 of SURVEY.md section 8):
oracle/_ref/StokesBEM = examples/StokesBEM.cpp with its FGMRES (examples/BEM/GMRES_Stokes.hpp:170-300), the
block-diagonal preconditioner (-diagonal: BlockDiagonalPC over a block_diagonal plan, include/executor/
EvalDiagonalSparse.hpp:12-80) and the local inner solver (-local: LocalPC over a local_evaluation plan, include/executor/
EvalLocalSparse.hpp:12-124), run on ONE thread in the build container.  The shipped examples/LaplaceBEM.cpp compiles its
FGMRES / local branches out (#if 1 ... #else, :285-317): with -local or -fgmres it solves nothing, which is recorded too.
usage: python tests/golden/make_precond_golden.py   (writes tests/golden/precond_lines.json)"""
import json
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")


def run(exe, args):
    with tempfile.TemporaryDirectory() as tmp:
        out = subprocess.check_output([os.path.join(REF, exe)] + args, cwd=tmp,
                                      env=dict(os.environ, OMP_NUM_THREADS="1")).decode()
    its = [(int(a), float(b), int(c)) for a, b, c in re.findall(r"it: (\d+), res: ([0-9.eE+-]+), fmm_req_p: (\d+)", out)]
    m = re.search(r"Final residual: ([0-9.eE+-]+), after (\d+) iterations", out)
    rec = {"args": args, "iterations": its, "final_residual": float(m.group(1)) if m else None,
           "n_iterations": int(m.group(2)) if m else 0,
           "solver_line": (re.search(r"Solver: .*", out) or [None])[0] if re.search(r"Solver: .*", out) else None}
    for key, pat in (("fx", r"Fx: ([0-9.eE+-]+)"), ("rhs_error", r"rhs error: ([0-9.eE+-]+)"),
                     ("relative_error", r"relative error: ([0-9.eE+-]+)")):
        mm = re.search(pat, out)
        if mm:
            rec[key] = float(mm.group(1))
    return rec


def main():
    out = {"stokes": [], "laplace": []}
    for rec in (4, 5):
        for flag in ("-fgmres", "-diagonal", "-local"):
            out["stokes"].append(run("StokesBEM", ["-recursions", str(rec), "-p", "8", "-k", "4", "-solver_tol", "1e-5", flag]))
    for flag in ("-local", "-fgmres"):
        out["laplace"].append(run("LaplaceBEM", ["-recursions", "4", "-p", "8", "-k", "4", "-solver_tol", "1e-6", flag]))
    path = os.path.join(ROOT, "tests", "golden", "precond_lines.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
