# cfg5 / cfg4 shard 0 of 8 on one GPU (development aid): bash scripts/gpu_one5.sh [lib ...]
for lib in "${@:-main}"; do
  if [ $lib != main ]; then export NT_LIB_PATH=$PWD/nettracer_b200/variants/libnt_$lib.so; else unset NT_LIB_PATH; fi
  echo "== $lib"
  NT_ONE_SHARDS=8 timeout 300 python scripts/gpu_one.py cfg5_mesh1m_8k_16spp_d5 f64 2 2>&1 | tail -1 | cut -c1-420
  NT_ONE_SHARDS=8 timeout 300 python scripts/gpu_one.py cfg5_mesh1m_8k_16spp_d5 f32 2 2>&1 | tail -1 | cut -c1-420
done
