V=chest-x-ray-vit_b200/csrc/build/variants
for n in default s6b2 s5b2 s4b4; do
  if [ $n = default ]; then unset VITK_LIB; else export VITK_LIB=$PWD/$V/libvitk_$n.so; fi
  echo "=== $n"; python tools/bench_gemm.py --quick 2>&1 | tail -14 | cut -c1-62
done
unset VITK_LIB
python tools/gemm_timeline.py "qkv fwd" "fc1 fwd" "out fwd" "fc2 wgrad" 2>&1 | grep -v "request→full by"
