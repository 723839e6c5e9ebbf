#!/bin/bash
# Builds libellcuda's objects at <commit> in a scratch worktree and compares their SASS with the current tree's, object by
# object, after removing what depends on the checkout path (anonymous-namespace symbol hashes) — the check behind "moving the
# host launchers behind #ifdef __CUDACC__ / the device code into headers did not change the kernels".
# usage: tools/sass_identity.sh <commit>     (run from the repo root; needs nvcc, no GPU)
set -e
ref=${1:?commit}
root=$(pwd)
wt=$(mktemp -d)/wt
git worktree add -q "$wt" "$ref"
(cd "$wt/ellalgo_rs_b200/csrc" && make -j8 > /dev/null 2>&1)
(cd ellalgo_rs_b200/csrc && make -j8 > /dev/null 2>&1)
norm() { cuobjdump -sass "$1" | grep -v "^Fatbin\|^$\|Function :\|\.headerflags\|^\s*code for\|^\s*\.section" | sed 's/_GLOBAL__N__[0-9a-f]*_[0-9]*_[a-z_0-9]*_cu_[0-9a-f]*/ANON/g'; }
regs() { sed 's/\/\* 0x[0-9a-f]* \*\///; s/\bU\?R[0-9]\+\b/R/g; s/\bU\?P[0-9]\b/P/g; s/  */ /g; s/ *$//' | grep -v '^ *$'; }   # mnemonics and operands without register numbers or encodings
echo "SASS of $(git rev-parse --short HEAD) against $(git rev-parse --short "$ref"), per object:"
for o in "$wt"/ellalgo_rs_b200/csrc/_obj/*.o; do
    f=$(basename "$o")
    cur="$root/ellalgo_rs_b200/csrc/_obj/$f"
    if [ ! -f "$cur" ]; then echo "  $f: not in the current tree"; continue; fi
    if cmp -s <(norm "$o") <(norm "$cur"); then echo "  $f: identical"
    elif cmp -s <(norm "$o" | regs) <(norm "$cur" | regs); then echo "  $f: identical up to register names ($(diff <(norm "$o") <(norm "$cur") | grep -c '^<') instructions renamed)"
    else echo "  $f: DIFFERENT ($(diff <(norm "$o") <(norm "$cur") | grep -c '^<') lines)"; fi
done
git worktree remove --force "$wt"
