#!/bin/bash
# Build the product library of another git revision as an experiment build (A/B runs on one GPU box):
#   bash tools/dev/build_rev.sh <rev> <name> [DEFS]  ->  saprobe-alac_b200/libalacb200_<name>.so
set -e
REV=$1; NAME=$2; ROOT=$(cd "$(dirname "$0")/../.." && pwd); T=$(mktemp -d)
git -C "$ROOT" archive "$REV" saprobe-alac_b200/csrc saprobe-alac_b200/host include | tar -x -C "$T"
make -C "$T/saprobe-alac_b200/csrc" variant NAME="$NAME" DEFS="$3" >/dev/null 2>&1
cp "$T/saprobe-alac_b200/libalacb200_$NAME.so" "$ROOT/saprobe-alac_b200/"
rm -rf "$T"; ls -la "$ROOT/saprobe-alac_b200/libalacb200_$NAME.so"
