#!/bin/sh
# oracle/ref_move.sh -- TEST INFRASTRUCTURE.  The reference's volume-data move, verbatim command
# string from moveVolumeData (reference utils/copy.go:116) with the helper container's bind
# paths /root/src and /root/dest replaced by host paths:
#   "find /root/src/ -maxdepth 1 -type f | xargs mv --target-directory=/root/dest; mv /root/src/* /root/dest"
# Put SRC and DST on different mounts (e.g. /dev/shm and /tmp) to reproduce the reference's
# cross-bind-mount EXDEV copy+unlink behaviour.
# Usage: ref_move.sh SRC DST
set -u
SRC="$1"; DST="$2"
exec sh -c "find $SRC/ -maxdepth 1 -type f | xargs mv --target-directory=$DST; mv $SRC/* $DST"
