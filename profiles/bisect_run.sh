# run one fuzz shape (FUZZ_ONLY) on several builds of the library: glow-tts-train_b200/libmas_ab_<tag>.so
cp glow-tts-train_b200/libmas_b200.so /tmp/libmas_keep.so
for v in "$@"; do
  cp glow-tts-train_b200/libmas_ab_$v.so glow-tts-train_b200/libmas_b200.so
  FUZZ_VERBOSE=1 timeout 300 python profiles/fuzz_fused.py 1000 > gpurun_out/fz_$v.log 2>&1
  echo "$v: $(grep -c 'ok$' gpurun_out/fz_$v.log) ok lines; $(tail -1 gpurun_out/fz_$v.log | cut -c1-120)"
done
cp /tmp/libmas_keep.so glow-tts-train_b200/libmas_b200.so
