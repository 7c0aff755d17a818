#!/usr/bin/env python
"""Where the fixtures in this directory come from (run in the build container, where
/root/reference is mounted; the GPU box never sees the reference tree).

  sir_incidence.csv         <- inst/sir_incidence.csv          (columns: cases, day)
  nested_sir_incidence.csv  <- inst/nested_sir_incidence.csv   (columns: cases, day, population)

They are the reference's own example DATA (the series its vignettes and BASELINE.json's
configs C1-C5 run on), copied byte for byte -- no reference source code is copied.
Everything else the tests pin (scale_log_weights cases, particle_filter_data layouts,
early-exit rules, reflect_proposal cases, the Kalman recursion) is written into the test
that uses it, next to the reference file:line it comes from.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/inst"
FILES = ["sir_incidence.csv", "nested_sir_incidence.csv"]

if __name__ == "__main__":
    check = "--check" in sys.argv
    for name in FILES:
        src, dst = os.path.join(SRC, name), os.path.join(HERE, name)
        if check:
            assert filecmp.cmp(src, dst, shallow=False), name + " differs from the reference's"
            print("ok", name)
        else:
            shutil.copyfile(src, dst)
            print("copied", name)
