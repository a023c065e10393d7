#!/bin/sh
# The known-answer fixtures of the reference's per-solver tests (testsuite.at:312-346):
# three 5x5 systems, one right-hand side file and six expected-answer files.  They are data,
# not source; copied verbatim from /root/reference/tests by this script.
set -e
cd "$(dirname "$0")"
for f in unsym.mm sym.mm sym_posdef.mm rhs1.mm unsym-default-ans.mm unsym-rhs1-ans.mm sym-default-ans.mm sym-rhs1-ans.mm sym_posdef-default-ans.mm sym_posdef-rhs1-ans.mm; do
  cp /root/reference/tests/$f .
done
