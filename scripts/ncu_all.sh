# ncu --set full captures of isolated launches (scripts/ncu_targets.py <target>); reports -> gpurun_out/ncu_<target>.ncu-rep
# usage: bash scripts/ncu_all.sh [targets...]   (default: conv geglu res320 ffdown attn tattn norms)
set -x
T=${@:-conv geglu res320 ffdown attn tattn norms}
for t in $T; do
  k="regex:igemm"; s=2
  case $t in ff|ff1|lnqkv) k="regex:ff_kernel";; attn) k="regex:attn2";; tattn) k="regex:attn_kernel";; norms) k="regex:gn_|layernorm|axpby_gn"; s=5;; esac
  c=1; [ $t = norms ] && c=5
  [ $t = sk ] && { s=2; c=4; }  # last linear launch + the three conv launches
  timeout 300 ncu --set full --clock-control none --import-source on -k $k -s $s -c $c -o gpurun_out/ncu_$t -f python scripts/ncu_targets.py $t > gpurun_out/ncu_$t.log 2>&1
done
ls -la gpurun_out/ncu_*.ncu-rep
