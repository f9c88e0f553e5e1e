"""optimize_stencil.py's main() in a FRESH process (what a user of the command-line tool waits for), timed from the
first line of this script: imports, tables, worker start, CUDA context, scan, high-resolution polish.

    python tools/cli_timing.py --div-free --dim 3 --dtrange 0.6718 1.0 --N 64 --Ngrid_low 128 --Ngrid_high 128
"""
import time
T0 = time.perf_counter()
import contextlib, hashlib, io, os, sys  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    from optimize_stencil_b200.cli import optimize_parser
    from optimize_stencil_b200.extended_stencil import Optimize
    t_import = time.perf_counter() - T0
    args = optimize_parser().parse_args(sys.argv[1:])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        opt = Optimize(args)
        x, f = opt.optimize()
    wall = time.perf_counter() - T0
    print("wall %.3f s (imports %.2f s)  norm %r  x %s  stdout sha1 %s\n  scan %s"
          % (wall, t_import, float(f), list(map(float, x)), hashlib.sha1(buf.getvalue().encode()).hexdigest()[:12], opt.scan_stats))
