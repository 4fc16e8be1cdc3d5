#!/usr/bin/env python3
"""Acceptance checker for d2q9-bgk results -- same command line, same verdict and same
report lines as the reference's check/check.py (check/check.py:1-151 there), written
afresh so that it can travel with this repository (tests/test_check_script.py runs
both on the same files where the reference is available and compares the output).

    check.py --ref-av-vels-file R1 --ref-final-state-file R2 \
             --av-vels-file A --final-state-file F [--tolerance PCT]

What is compared (reference check.py:57-63, 83-99, 136-151):
  * av_vels.dat column 1, every step;
  * final_state.dat columns 0, 1 (coordinates: must be identical, in identical order)
    and 5 (pressure);
  * percentage difference 100 * (ref - sim) / sim, worst absolute value per file;
  * FAIL if that worst value is not finite or exceeds the tolerance (default 1 %).
Exit status 0 = both passed, 1 = anything else.
"""
import argparse
import sys

import numpy as np


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="Testing script for HPC LBM coursework",
                                 fromfile_prefix_chars="@",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--tolerance", nargs=1, type=float, default=[1],
                    help="Percentage tolerance to match against reference results")
    for opt, text in (("--ref-av-vels-file", "reference av_vels results file"),
                      ("--ref-final-state-file", "reference final_state results file"),
                      ("--av-vels-file", "calculated av_vels results file"),
                      ("--final-state-file", "calculated final_state results file")):
        ap.add_argument(opt, nargs=1, required=True, help=text)
    return ap.parse_args(argv)


def read_pair(av_vels_path, final_state_path):
    av = np.loadtxt(av_vels_path, usecols=[1], ndmin=1)
    fs = np.loadtxt(final_state_path, usecols=[0, 1, 5], ndmin=2)
    return av, fs


def worst_difference(ref, sim):
    """Largest |percentage difference| and where it is."""
    delta = ref - sim
    with np.errstate(divide="ignore", invalid="ignore"):
        pct = 100.0 * (delta / (ref - delta))
    k = int(np.argmax(np.abs(pct)))
    return {"where": k, "delta": delta[k], "pct": pct[k], "sim": sim[k], "ref": ref[k],
            "total": float(np.sum(np.abs(delta)))}


def failed(d, tolerance):
    return (not np.isfinite(d["pct"])) or abs(d["pct"]) > tolerance


def main(argv=None):
    a = parse_args(argv)
    tol = a.tolerance[0]
    av_ref, fs_ref = read_pair(a.ref_av_vels_file[0], a.ref_final_state_file[0])
    av_sim, fs_sim = read_pair(a.av_vels_file[0], a.final_state_file[0])

    if fs_ref.shape != fs_sim.shape or np.any(fs_ref[:, 0:2] != fs_sim[:, 0:2]):
        print("Final state files coordinates were not the same")
        return 1
    if av_ref.size != av_sim.size:
        print("Different number of steps in av_vels files")
        return 1

    d_av = worst_difference(av_ref, av_sim)
    print("Total difference in av_vels : {:.12E}".format(d_av["total"]))
    print("Biggest difference (at step {:d}) : {:.12E}".format(d_av["where"], d_av["delta"]))
    print("  {:.12E} vs. {:.12E} = {:.2g}%".format(d_av["sim"], d_av["ref"], d_av["pct"]))
    print()

    d_fs = worst_difference(fs_ref[:, 2], fs_sim[:, 2])
    k = d_fs["where"]
    print("Total difference in final_state : {:.12E}".format(d_fs["total"]))
    print("Biggest difference (at coord ({:d},{:d})) : {:.12E}".format(
        int(fs_sim[k, 0]), int(fs_sim[k, 1]), d_fs["delta"]))
    print("  {:.12E} vs. {:.12E} = {:.2g}%".format(d_fs["sim"], d_fs["ref"], d_fs["pct"]))
    print()

    bad_fs, bad_av = failed(d_fs, tol), failed(d_av, tol)
    if bad_fs:
        print("final state failed check")
    if bad_av:
        print("av_vels failed check")
    if bad_fs or bad_av:
        return 1
    print("Both tests passed!")
    return 0


if __name__ == "__main__":
    sys.exit(main())
