# ncu --set full of the two weight-gradient forms on two layers (raw CSV export, small)
O=gpurun_out
for L in down1.c2 down3.c2; do
  timeout 200 ncu --set full --clock-control none -k regex:wgrad_umma -s 1 -c 1 -f -o $O/r01_full_wgrad1_$L python tools/prof_wgrad.py $L > $O/wg1_$L.log 2>&1
  timeout 200 ncu --set full --clock-control none -k regex:wgrad2_umma -s 1 -c 1 -f -o $O/r01_full_wgrad2_$L python tools/prof_wgrad.py $L > $O/wg2_$L.log 2>&1
  for V in 1 2; do ncu -i $O/r01_full_wgrad${V}_$L.ncu-rep --page raw --csv > $O/r01_full_wgrad${V}_$L.raw.csv 2>/dev/null && rm -f $O/r01_full_wgrad${V}_$L.ncu-rep; done
done
du -sh $O
