"""example.py of the reference on the B200 ops: 2D Z_K gauge theory, (flavour HOTRG +) ATRG / TRG.

Same command line and the same printed columns as the reference's example.py (reference example.py:60-196):

    python examples/example.py                       # ATRG, Dcutxy 32, 5 steps, anti-periodic
    python examples/example.py --trg --Dcutxy 64 --cgsteps 10
    python examples/example.py --block               # = the reference's example_block.py
    python examples/example.py --log run.jsonl --checkpoint ckpt/ --resume

The initial-tensor construction (gauge2d.tensor_preparation: sympy + the `sparse` package, 18 minutes on a
CPU) is outside the accelerated path: the site tensor comes from a file written once by
the reference (tests/golden/make_z2_tensor.py).  The committed fixture is the reference's default parameter set
(Z_2, N_f=1, beta=m=q=a=1, mu=0); other parameters need `--tensor file.npz` made the same way.
"""
import argparse
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def time_display(dt):
    return "%.3g s" % dt if dt < 60 else "%d min %.0f s" % (dt // 60, dt % 60)


def main(argv=None, force_block=False):
    parse = argparse.ArgumentParser()
    parse.add_argument('--beta', default=1.0, type=float)
    parse.add_argument('--mass', default=1.0, type=float)
    parse.add_argument('--charge', default=1.0, type=float)
    parse.add_argument('--spacing', default=1.0, type=float)
    parse.add_argument('--mu', default=0.0, type=float)
    parse.add_argument('--Nf', default=1, type=int)
    parse.add_argument('--K', default=2, type=int)
    parse.add_argument('--cgsteps', default=5, type=int)
    parse.add_argument('--Dcutz', default=32, type=int)
    parse.add_argument('--Dcutxy', default=32, type=int)
    parse.add_argument('--boundary_conditions', default="anti-periodic")
    parse.add_argument('--trg', '--TRG', dest="trg", default=False, action='store_true')
    parse.add_argument('--block', default=force_block, action='store_true', help="block format (example_block.py)")
    parse.add_argument('--tensor', default=None, help=".npz with the initial site tensor (data, statistics, encoder, format)")
    parse.add_argument('--from-ab', dest="from_ab", default=False, action='store_true',
                       help="build the initial tensor on the GPU from the A / B tensors (the compression stages of the "
                            "reference's tensor_preparation: gauge2d.tensor_from_AB) instead of loading the compressed fixture")
    parse.add_argument('--log', default=None, help="JSON-lines run log")
    parse.add_argument('--checkpoint', default=None, help="directory for per-step tensor checkpoints")
    parse.add_argument('--resume', default=False, action='store_true')
    args = parse.parse_args(argv)

    import grassmanntn_b200 as gtn
    gauge = gtn.gauge2d_block if args.block else gtn.gauge2d
    default = (args.beta, args.mass, args.charge, args.spacing, args.mu, args.K) == (1.0, 1.0, 1.0, 1.0, 0.0, 2)
    if args.tensor is None and not (default and args.Nf in (1, 2)):
        sys.exit("only the default parameter set has a committed initial tensor; build one with the reference's "
                 "tensor_preparation (tests/golden/make_z2_tensor.py) and pass it with --tensor")
    bc = args.boundary_conditions
    print(" parameters: β=%s, m=%s, μ=%s, q=%s, a=%s, Nf=%d, Z_%d, Dcutz=%d, Dcutxy=%d, %s, %s%s"
          % (args.beta, args.mass, args.mu, args.charge, args.spacing, args.Nf, args.K, args.Dcutz, args.Dcutxy, bc,
             "trg" if args.trg else "atrg", ", block" if args.block else ""))
    t0 = time.time()
    if args.from_ab:
        T, terr = gauge.tensor_from_AB(*gauge.load_AB_tensors())
        print(" compression error:", '{:.3g}'.format(float(terr)))
    else:
        T = gauge.load_initial_tensor(args.tensor)
    if args.block:
        T = T.toblock()
    shp = (lambda X: X.effective_shape) if args.block else (lambda X: X.shape)
    print(" initial_tensor:", tuple(shp(T)[:2]), "   ", time_display(time.time() - t0))

    logNorm = 0.0
    for i in range(int(math.log2(args.Nf))):
        t0 = time.time()
        T, Tnorm, err = gauge.hotrg3dz(T, T, args.Dcutz, intermediate_dcut=args.Dcutz, iternum=i, error_test=True)
        logNorm = 2 * logNorm + np.log(Tnorm)
        print(" flavor_cg:", tuple(shp(T)[:2]), "   ", '{:.3g}'.format(float(err)), "   ", time_display(time.time() - t0))
    T = gauge.zcap(T)

    t0 = time.time()
    T, records = gauge.coarse_grain(T, cgsteps=args.cgsteps, dcut=args.Dcutxy, method="trg" if args.trg else "atrg",
                                    boundary_conditions=bc, error_test=True, log=args.log, logNorm0=logNorm,
                                    checkpoint_dir=args.checkpoint, resume=args.resume)
    head = (args.beta, args.mass, args.mu, args.charge, args.spacing, args.Nf, args.K)
    for r in records:
        F = r["F"]
        name = "   _ini" if r["Tnorm"] is None else ("   _trg" if args.trg else "  _atrg")
        tail = () if r["err"] is None else ("   ", '{:.3g}'.format(r["err"]))
        print(name + ":", r["vol"], *head, "   ", F.real, "   ", F.imag, "   ", tuple(r["shape"]), *tail)
    print(" total:", time_display(time.time() - t0))
    return records


if __name__ == "__main__":
    main()
