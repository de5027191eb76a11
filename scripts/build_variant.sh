#!/bin/bash
# build_variant.sh <git-ref|WORKTREE> <name>: compiles the csrc of a git ref (or the working tree) into
# v-diffusion-torch_b200/lib/libvdt_b200_<name>.so for same-box A/B runs (VDT_LIB=...)
set -e
ref=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
if [ "$ref" = "WORKTREE" ]; then cp -r "$root/v-diffusion-torch_b200/csrc" "$tmp/csrc"; mkdir -p "$tmp/inc"; cp "$root/include/vdt_b200.h" "$tmp/inc/";
else mkdir -p "$tmp/csrc" "$tmp/inc"; for f in $(git -C "$root" ls-tree --name-only "$ref" v-diffusion-torch_b200/csrc/); do git -C "$root" show "$ref:$f" > "$tmp/csrc/$(basename $f)"; done; git -C "$root" show "$ref:include/vdt_b200.h" > "$tmp/inc/vdt_b200.h"; fi
sed -i 's#"../../include/vdt_b200.h"#"../inc/vdt_b200.h"#' "$tmp/csrc/plan.cu"
nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o "$root/v-diffusion-torch_b200/lib/libvdt_b200_$name.so" "$tmp"/csrc/*.cu
rm -rf "$tmp"; echo "built $name from $ref"
