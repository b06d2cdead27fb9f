#!/bin/bash
# A/B of two builds of the library on the SAME box: tools/ab_lib.sh <other .so>  (CL_B200_LIB selects the library)
echo "== default build"; python tools/ab_env.py ""
echo "== $1"; CL_B200_LIB=$(realpath "$1") python tools/ab_env.py ""
