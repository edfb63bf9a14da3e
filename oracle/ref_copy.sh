#!/bin/sh
# oracle/ref_copy.sh -- TEST INFRASTRUCTURE.  The reference's own copy engine, verbatim:
#   cpRFPOption = "(cd %s; tar c .) | (cd %s; tar x)"          (reference utils/copy.go:17-19)
#   CopyDir(src,dest) runs it through sh -c                      (reference utils/copy.go:21-27)
# Usage: ref_copy.sh SRC DST      (DST must exist, as the reference assumes)
set -u
SRC="$1"; DST="$2"
exec sh -c "(cd $SRC; tar c .) | (cd $DST; tar x)"
