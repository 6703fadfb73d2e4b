#!/usr/bin/env bash
# Build a library VARIANT for A/B runs on one GPU box:  scripts/ab_variant.sh <name> <git-ref> <csrc file>...
# The variant is the current rdo_ptq_b200/csrc tree with the listed files taken from <git-ref>; it lands in
# build/variants/<name>/libb200lic.so (build/ is git-ignored but travels with gpurun) and is selected with
# B200LIC_LIB=<path> (rdo_ptq_b200/_lib.py).  Same ABI as the in-tree library or the run fails loudly.
set -euo pipefail
name=$1; ref=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p "$tmp/rdo_ptq_b200/csrc" "$tmp/include" "$root/build/variants/$name"
cp "$root"/rdo_ptq_b200/csrc/* "$tmp/rdo_ptq_b200/csrc/"
cp "$root/include/b200lic.h" "$tmp/include/"
for f in "$@"; do git -C "$root" show "$ref:rdo_ptq_b200/csrc/$f" > "$tmp/rdo_ptq_b200/csrc/$f"; done
make -C "$tmp/rdo_ptq_b200/csrc" -j8 > "$tmp/make.log" 2>&1 || { tail -20 "$tmp/make.log"; exit 1; }
cp "$tmp/rdo_ptq_b200/lib/libb200lic.so" "$root/build/variants/$name/libb200lic.so"
rm -rf "$tmp"
echo "build/variants/$name/libb200lic.so"
