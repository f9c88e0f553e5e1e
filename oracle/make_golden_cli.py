"""TEST INFRASTRUCTURE ONLY -- captures the stdout of the UNMODIFIED reference command-line tools.

    python oracle/make_golden_cli.py        (build container only: needs /root/reference)

Writes tests/golden/cli_*.txt; tests/test_cli.py replays the same command lines through
optimize_stencil_b200.cli and compares line by line (numbers to tolerance, text exactly)."""
import contextlib
import io
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
RUNS = {
    "cli_calc_lehe2d": ("calculate_omega.py", ["--lehe", "--params", "0.96"]),
    "cli_calc_lehe3d": ("calculate_omega.py", ["--lehe", "--params", "0.96", "--dim", "3", "--Ngrid_high", "48"]),
    "cli_calc_yee_epoch": ("calculate_omega.py", ["--yee", "--dim", "3", "--Ngrid_high", "32", "--Y", "1.5", "--output", "epoch"]),
    "cli_calc_pukhov_array": ("calculate_omega.py", ["--pukhov", "--Ngrid_high", "64", "--output", "array", "--weight", "xaxis_soft", "--weight-params", "0.3"]),
    "cli_calc_laser2d": ("calculate_omega.py", ["--lehe", "--params", "0.9", "--Ngrid_high", "64", "--physical-dx", "0.05", "--laser-wavelength", "0.8", "--laser-direction", "1", "0.2"]),
    "cli_calc_laser3d": ("calculate_omega.py", ["--yee", "--dim", "3", "--Ngrid_high", "24", "--physical-dx", "0.05", "--laser-wavelength", "0.8"]),
    "cli_opt_divfree": ("optimize_stencil.py", ["--div-free", "--dtrange", "0.6718", "1.0", "--singlecore", "--Ngrid_low", "30", "--Ngrid_high", "60"]),
    # infeasible parameters: the grid minimum of sqrtarg is negative, stencil_ok / dt_ok print as NumPy scalars
    "cli_calc_infeasible2d": ("calculate_omega.py", ["--no-symmetric-beta", "--Ngrid_high", "40", "--params", "0.3", "0.3", "-0.3", "0.3", "0.7"]),
    "cli_calc_infeasible3d": ("calculate_omega.py", ["--dim", "3", "--Ngrid_high", "16", "--params", "0.4", "0.6", "-0.5", "-0.7", "0.24", "-0.9", "0.1"]),
    # --write-omega: the gnuplot table of dispersion.py:199-232, kept byte for byte as tests/golden/<name>.dat
    "omega_yee2d": ("calculate_omega.py", ["--yee", "--Ngrid_high", "12", "--write-omega", "@OUT@"]),
    "omega_lehe3d": ("calculate_omega.py", ["--lehe", "--params", "0.9", "--dim", "3", "--Ngrid_high", "9", "--write-omega", "@OUT@"]),
    "omega_soft2d": ("calculate_omega.py", ["--pukhov", "--Ngrid_high", "17", "--Y", "0.8", "--weight", "xaxis_soft", "--weight-params", "0.3", "--write-omega", "@OUT@"]),
}


def main():
    import numpy as np
    ref_shim.load_reference()
    # dispersion.py:231 hands np.vstack a map object, which NumPy >= 1.16 rejects: list() it from outside the
    # reference tree, for the --write-omega runs only
    plain_vstack = np.vstack
    only = set(sys.argv[1:])
    for name, (script, argv) in RUNS.items():
        if only and name not in only:
            continue
        dat = os.path.join(OUT, name + ".dat")
        argv = [dat if a == "@OUT@" else a for a in argv]
        np.vstack = (lambda tup, **kw: plain_vstack(list(tup), **kw)) if name.startswith("omega_") else plain_vstack
        buf = io.StringIO()
        old = sys.argv
        sys.argv = [script] + argv
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
        try:
            with contextlib.redirect_stdout(buf):
                try:
                    runpy.run_path(os.path.join(ref_shim.REFERENCE_ROOT, script), run_name="__main__")
                except SystemExit:
                    pass
        finally:
            sys.argv = old
            sys.path.remove(ref_shim.REFERENCE_ROOT)
        np.vstack = plain_vstack
        argv = ["@OUT@" if a == dat else a for a in argv]
        with open(os.path.join(OUT, name + ".txt"), "w") as f:
            f.write(" ".join(argv) + "\n")
            f.write(buf.getvalue())
        print(name, len(buf.getvalue().splitlines()), "lines")


if __name__ == "__main__":
    main()
