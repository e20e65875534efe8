# ncu --set full of the training-step elementwise kernels (small reports: few launches each, no source import)
B="python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1"
timeout 200 ncu --set full --clock-control none -k regex:bn_apply_kernel -c 2 -f -o gpurun_out/r01h_bn_apply $B > gpurun_out/r01h_a.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"bn_bwd_reduce|bn_bwd_apply" -c 4 -f -o gpurun_out/r01h_bn_bwd_dec $B > gpurun_out/r01h_b.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"bn_bwd_reduce|bn_bwd_apply" -s 32 -c 4 -f -o gpurun_out/r01h_bn_bwd_enc $B > gpurun_out/r01h_c.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"up_input_bwd|outconv" -c 4 -f -o gpurun_out/r01h_misc $B > gpurun_out/r01h_d.log 2>&1
ls -la gpurun_out
