"""Command-line tools: the reference's optimize_stencil.py and calculate_omega.py on the GPU objective.

Same flags, defaults and printed output as the reference scripts (optimize_stencil.py:32-117,
calculate_omega.py:30-184); the option tables below are data, the parsers are built from them.
"""
import argparse
import shlex
import sys

import numpy as np

from . import extended_stencil
from .extended_stencil import Dispersion2D, Dispersion3D, Optimize, get_stencil, weight_functions
from .extended_stencil.optimize import _trace

_AXES = ("x", "y", "z")
_BETAS = ("xy", "xz", "yx", "yz", "zx", "zy")

# (flags, kwargs) shared by both tools
_STENCIL_OPTS = [
    (("--dim",), dict(default=2, type=int, choices=[2, 3],
                      help="Dimension of the stencil to be optimized (default: %(default)s).")),
    (("--div-free",), dict(action="store_true",
                           help="Constrain the derivative of div B == 0 (default: %(default)s).")),
    (("--symmetric-axes",), dict(type=int, choices=[0, 1, 2], default=0,
                                 help="Define the count of axes of the stencil, that should be identical to "
                                      "another axis (default: %(default)s).")),
    (("--symmetric",), dict(dest="symmetric_axes", action="store_const", const=1,
                            help="Make sure the stencil has 1 symmetric axis, equivalent to --symmetric-axes 1")),
    (("--no-symmetric",), dict(dest="symmetric_axes", action="store_const", const=0,
                               help="Make sure all axes of the stencil are free, equivalent to --symmetric-axes 0")),
    (("--no-symmetric-beta",), dict(dest="symmetric_beta", action="store_false",
                                    help="Do not demand that the beta coefficients should form a symmetric matrix. "
                                         "Warning: this will introduce spurious dimensions into the optimization "
                                         "problem and thus lead to unreproducible results.")),
    (("--Y",), dict(default=1, type=float, help="Grid aspect ratio dy/dx (default: %(default)s).")),
    (("--Z",), dict(default=1, type=float, help="Grid aspect ratio dz/dx (default: %(default)s).")),
    (("--dt-multiplier",), dict(default=1.0, type=float,
                                help="Multiplier to be applied to time step as returned by CFL condition "
                                     "(default: %(default)s).")),
]
_WEIGHT_OPTS = [
    (("--weight",), dict(choices=["equal", "xaxis", "xaxis_soft"], default="equal",
                         help="Choose weight function (default: %(default)s).")),
    (("--weight-params",), dict(nargs="*", type=float, help="Parameters for weight function.", default=[])),
    (("--write-omega",), dict(default=None, help="Write omega to file OUTFILE", metavar="OUTFILE")),
]
_DESCRIPTION = "This script calculates the optimal coefficients for a FDTD stencil."


def _add(parser, table):
    for flags, kw in table:
        parser.add_argument(*flags, **kw)


def _range_opts():
    out = []
    for ax in _AXES:
        out.append((("--delta%srange" % ax,), dict(default=[-1, 0.25], nargs=2, type=float, metavar=("min", "max"),
                                                   help="Range of delta%s (default: %%(default)s)." % ax)))
    for b in _BETAS:
        out.append((("--beta%srange" % b,), dict(default=[-1, 1], nargs=2, type=float, metavar=("min", "max"),
                                                 help="Range of beta%s (default: %%(default)s)." % b)))
    out.append((("--dtrange",), dict(default=[0.1, 1], nargs=2, type=float, metavar=("min", "max"),
                                     help="Range of dt in units of dx (default: %(default)s).")))
    return out


def optimize_parser():
    p = argparse.ArgumentParser(description=_DESCRIPTION)
    p.add_argument("--N", default=3, type=int,
                   help="Number of scan points for initial conditions in each coefficient (default: %(default)s).")
    p.add_argument("--no-scan-dt", action="store_false", dest="scan_dt",
                   help="Do not scan through the allowed range of the time step for initial conditions and only use "
                        "one starting value of 95%% of the CFL condition (default: %(default)s).")
    p.set_defaults(scan_dt=True)
    p.add_argument("--Ngrid_low", default=50, type=int,
                   help="Gridsize for the fast calculation of the norm in the simplex algorithm (default: %(default)s).")
    p.add_argument("--Ngrid_high", default=200, type=int,
                   help="Gridsize for the accurate norm calculation after simplex  run (default: %(default)s).")
    _add(p, _STENCIL_OPTS)
    _add(p, _range_opts())
    _add(p, _WEIGHT_OPTS[:2])
    p.add_argument("--singlecore", action="store_true",
                   help="Accepted for compatibility; the GPU build has no process pool.")
    _add(p, _WEIGHT_OPTS[2:])
    p.add_argument("-v", "--version", action="store_true")
    return p


def main_optimize(argv=None):
    args = optimize_parser().parse_args(argv)
    if args.version:
        print('optimize_stencil.py {}'.format(extended_stencil.__version__))
        return 0
    print('Arguments are:', args)
    print('Starting Optimization.')
    opt = Optimize(args)
    _trace('Optimize object built')
    x, fmin = opt.optimize()
    coefficients = opt.stencil.coefficients(x)
    print('\nOptimization finished. Results:\n')
    print("norm=", fmin)
    print('stencil_ok=', opt.dispersion_high.stencil_ok(x))
    print('dt_ok=', opt.dispersion_high.dt_ok(x))
    print("\n\n\n * Please include the following in your input.deck: *\n\n")
    for line in ("begin:control\n", "    [ ... your regular control block content ... ]\n",
                 "    maxwell_solver = custom", "    dt_multiplier = 1\n", "end:control\n\n", "begin:stencil\n",
                 "    # These coefficients have been calculated with",
                 "    # optimize_stencil.py version {}".format(extended_stencil.__version__), "",
                 "    # using the command line",
                 "    # \"optimize_stencil.py {}\"".format(" ".join(map(shlex.quote, sys.argv[1:] if argv is None else argv))),
                 ""):
        print(line)
    for name, value in zip(opt.stencil.Coefficients._fields, coefficients[0]):
        if name.startswith('alpha'):
            continue
        print(('    {} = {} * dx / c' if name == "dt" else '    {} = {}').format(name, value))
    print()
    print("end:stencil\n")
    if args.write_omega:
        opt.omega_output(args.write_omega, x)
    _trace('output written')
    return 0


def calculate_parser():
    p = argparse.ArgumentParser(description=_DESCRIPTION)
    p.add_argument("--Ngrid_high", default=200, type=int,
                   help="Gridsize for the accurate norm calculation after simplex  run (default: %(default)s).")
    _add(p, _STENCIL_OPTS)
    p.add_argument("--params", nargs="*", type=float, help="Parameters for stencil.")
    for preset in ("lehe", "yee", "pukhov"):
        p.add_argument("--" + preset, action="store_const", dest="type", const=preset)
    p.set_defaults(type="free")
    _add(p, _WEIGHT_OPTS[:2])
    p.add_argument("--output", default="standard", choices=["standard", "array", "epoch"],
                   help="Output format, 'standard' prints a list of the named coefficients, 'array' prints the "
                        "returned array x as it is with [dt, beta{xyz}{xyz}, delta{xyz}, norm] and 'epoch' prints it "
                        "compatible with the input decks of the EPOCH-Code.")
    _add(p, _WEIGHT_OPTS[2:])
    p.add_argument("--physical-dx", type=float, default=None,
                   help="Physical dx value for group/phase velocity calculations")
    p.add_argument("--laser-wavelength", type=float, default=None,
                   help="Calculate phase and group velocity for given wavelength")
    p.add_argument("--laser-direction", nargs="*", type=float, default=None,
                   help="Point laser in this direction (vector normalized automatically)")
    return p


def _preset_parameters(args, stencil, dispersion):
    """calculate_omega.py:88-120: free / yee / pukhov / lehe starting vectors."""
    names = stencil.Parameters._fields
    if args.type == "free":
        return args.params
    x = [0] * len(names)
    if args.type in ("pukhov", "lehe"):
        scale = dict(x=1.0, y=args.Y ** 2, z=args.Z ** 2)
        for i, name in enumerate(names):
            if name.startswith('beta'):
                x[i] = 0.125 / scale[name[5]]
    if args.type == "lehe":
        x[0] = args.params[0]
        for i, name in enumerate(names):
            if name == 'deltax':
                x[i] = 0.25 * (1 - 1 / x[0] ** 2 * np.sin(np.pi * x[0] / 2.0))
    else:
        x[0] = np.asarray(dispersion.dt_ok(x)).item()
    return x


def main_calculate(argv=None):
    args = calculate_parser().parse_args(argv)
    print(args)
    if args.type in ('lehe', 'pukhov') and args.div_free:
        print('Stencil of type {}, setting option --no-div-free'.format(('lehe', 'pukhov')))
        args.div_free = False
    if args.type in ('lehe') and args.symmetric_axes > 0:
        print('Stencil of type {}, setting option --no-symmetric'.format(('lehe')))
        args.symmetric = False
    stencil = get_stencil(args)
    print('Selected stencil', stencil.__class__.__name__, 'requires parameters', stencil.Parameters._fields)
    if args.dim == 2:
        dispersion = Dispersion2D(args.Y, args.Ngrid_high, stencil)
    else:
        dispersion = Dispersion3D(args.Y, args.Z, args.Ngrid_high, stencil)
    if args.weight in weight_functions:
        weight_functions[args.weight](*args.weight_params)(dispersion)
    dispersion.dt_multiplier = args.dt_multiplier
    x = _preset_parameters(args, stencil, dispersion)
    print('x=', x)
    print('stencil_ok=', dispersion.stencil_ok(x))
    print('dt_ok=', dispersion.dt_ok(x))
    fmin = dispersion.norm(x)
    coefficients = stencil.coefficients(x)
    pairs = list(zip(stencil.Coefficients._fields, coefficients[0]))
    if args.output == 'standard':
        print("norm=", fmin, "\n")
        for item in pairs:
            print('{}={}'.format(*item))
    if args.output == 'array':
        print([*coefficients[0], fmin])
    if args.output == 'epoch':
        print("norm=", fmin, "\n")
        print("\tmaxwell_solver=free")
        for item in pairs:
            print('\tstencil_{}={}'.format(*item))

    if args.laser_wavelength:
        if not args.physical_dx:
            print("To calculate dispersion at specified laser wavelength/k vector please give --physical-dx")
            return 1
        kl = np.zeros(args.dim)
        if args.laser_direction:
            kl[:] = args.laser_direction
            kl /= np.sqrt(np.sum(kl * kl))
        else:
            kl[0] = 1
        kl *= 2 * np.pi / args.laser_wavelength * args.physical_dx
        kabs = np.sqrt(np.sum(kl * kl))
        if args.dim == 2:
            spline = dispersion.omega_spline(x)
            vph = spline(*kl) / kabs
            vg = np.sqrt(spline(*kl, dx=1) ** 2 + spline(*kl, dy=1) ** 2)
        else:
            omega_interp, vg_interp = dispersion.omega_interp(x)
            vph = omega_interp(kl) / kabs
            vg = vg_interp(kl)
        print("Specified laser wavelength lambda={} dx leads to vph={} c and vg={} c.".format(
            args.laser_wavelength / args.physical_dx, np.asarray(vph).item(), np.asarray(vg).item()))
    if args.write_omega:
        dispersion.omega_output(args.write_omega, x)
    return 0


if __name__ == "__main__":
    tool = sys.argv[1] if len(sys.argv) > 1 else ""
    sys.exit({"optimize": main_optimize, "calculate": main_calculate}.get(
        tool, lambda a: print("usage: python -m optimize_stencil_b200.cli {optimize|calculate} [flags]") or 2)(sys.argv[2:]))
