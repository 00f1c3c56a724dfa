# Rebuilds libsnb.so with each combination of the packed-fp32 stage switches of the fused kernel
# (SNB_PACK_EW / SNB_PACK_FFT / SNB_PACK_UNPACK) and benches it: `gpurun -- bash tools/pack_variants.sh`
set -e
cd shennong_b200/csrc
for v in "0 0 0" "0 1 0" "0 1 1" "1 1 0" "1 1 1" "0 0 1" "1 0 0"; do
  set -- $v
  touch features.cu
  SNB_EXTRA_NVCC_FLAGS="-DSNB_PACK_EW=$1 -DSNB_PACK_FFT=$2 -DSNB_PACK_UNPACK=$3" bash build.sh > /dev/null 2>&1
  for d in 1.0 0.0; do
    (cd ../..; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --dither $d 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('EW=$1 FFT=$2 UN=$3 dither=$d', '%.4g' % j['value'], '%.3f' % j['roofline']['kernel_ms'])")
  done
done
# leave the default build behind
touch features.cu && bash build.sh > /dev/null 2>&1
