set -x
for t in conv geglu res320 attn tattn norms; do
  k="regex:igemm"; s=2
  case $t in attn) k="regex:attn2";; tattn) k="regex:attn_kernel";; norms) k="regex:gn_|layernorm"; s=3;; esac
  c=1; [ $t = norms ] && c=3
  timeout 300 ncu --set full --clock-control none --import-source on -k $k -s $s -c $c -o gpurun_out/ncu6_$t -f python scripts/ncu_targets.py $t > gpurun_out/ncu6_$t.log 2>&1
done
ls -la gpurun_out/ncu6_*.ncu-rep
