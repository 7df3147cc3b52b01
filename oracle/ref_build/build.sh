#!/bin/bash
# TEST INFRASTRUCTURE: compiles the reference's own FlowNet2 op kernels, UNMODIFIED and from where they lie under
# /root/reference, for sm_100a into oracle/_ref/libref_ops.so (git-ignored; it travels to the GPU box with the snapshot).
# The reference's build system (torch.utils.ffi + make.sh with -arch=sm_30/52) is not run; this is the whole recipe.
# Only the .cu kernel files are compiled; the THC-dependent C glue is replaced by oracle/ref_build/ref_harness.cu (ours).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${VECVAD_REFERENCE:-/root/reference}/FlowNet2_src/models/components/ops"
OUT="$HERE/../_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
if [ ! -d "$REF" ]; then
    echo "oracle/ref_build: $REF not present -- keeping any prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$OUT"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -Xcompiler -fPIC -I$HERE/shim"
$NVCC $FLAGS -I"$REF/correlation/src" -c "$REF/correlation/src/correlation_cuda_kernel.cu" -o "$OUT/correlation_cuda_kernel.o"
$NVCC $FLAGS -I"$REF/resample2d/src" -c "$REF/resample2d/src/Resample2d_kernel.cu" -o "$OUT/Resample2d_kernel.o"
$NVCC $FLAGS -I"$REF/channelnorm/src" -c "$REF/channelnorm/src/ChannelNorm_kernel.cu" -o "$OUT/ChannelNorm_kernel.o"
$NVCC $FLAGS -I"$REF/correlation/src" -I"$REF/resample2d/src" -I"$REF/channelnorm/src" -c "$HERE/ref_harness.cu" -o "$OUT/ref_harness.o"
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libref_ops.so" "$OUT/correlation_cuda_kernel.o" "$OUT/Resample2d_kernel.o" \
    "$OUT/ChannelNorm_kernel.o" "$OUT/ref_harness.o" -lcudart
echo "$OUT/libref_ops.so"
