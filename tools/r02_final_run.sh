# The round's closing measurement on one B200 (run through gpurun; everything lands in gpurun_out/, at most 64 MiB per call):
#   bash tools/r02_final_run.sh bench    GPU tests, bench line, reference arm, ncu launch list of the bench command
#   bash tools/r02_final_run.sh ncu      ncu --set full captures of bs23_kernel / dp54_kernel / mlp_tc_gemm_kernel as shipped
set -x
if [ "$1" = "bench" ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_all.log
python __graft_entry__.py smoke > gpurun_out/r02p_smoke.log 2>&1
python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
python bench.py --impl reference > gpurun_out/r02p_ref.json 2> gpurun_out/r02p_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02p_bench_launch_list.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/t_all.log gpurun_out/r02p_smoke.log; wc -c gpurun_out/r02p_bench.json gpurun_out/r02p_ref.json
else
export PFR_ATOL=1e-12 PFR_EOFF_METHOD=dp54 PFR_EOFF_TOL=1e-7
ncu --set full --clock-control none --import-source on -k regex:bs23_kernel -s 1 -c 1 -f -o gpurun_out/r02p_bs23 python tools/profile_target.py 524288 64 bs23 3e-7 > gpurun_out/ncu_bs23.log 2>&1
ncu --set full --clock-control none -k regex:dp54_kernel -s 1 -c 1 -f -o gpurun_out/r02p_dp54 python tools/profile_target.py 524288 64 bs23 3e-7 > gpurun_out/ncu_dp54.log 2>&1
ncu --set full --clock-control none -k regex:mlp_tc_gemm -s 8 -c 3 -f -o gpurun_out/r02p_mlp python tools/profile_target.py 151552 64 bs23 3e-7 > gpurun_out/ncu_mlp.log 2>&1
# the reports are ~26 MB each: digest them on the box and keep the digests (gpurun brings back at most 64 MiB per call)
python tools/ncu_summary.py gpurun_out/r02p_bs23.ncu-rep > gpurun_out/r02p_ncu_full_bs23_fp64.txt
ncu -i gpurun_out/r02p_bs23.ncu-rep --page source --csv > gpurun_out/r02p_bs23_source.csv 2>/dev/null
python tools/sass_mix.py gpurun_out/r02p_bs23_source.csv 30 > gpurun_out/r02p_bs23_executed_mix.txt
python tools/ncu_summary.py gpurun_out/r02p_dp54.ncu-rep > gpurun_out/r02p_ncu_full_dp54_fp64.txt
python tools/ncu_summary.py gpurun_out/r02p_mlp.ncu-rep > gpurun_out/r02p_ncu_full_mlp_tc_gemm_f16x3.txt
rm -f gpurun_out/r02p_dp54.ncu-rep gpurun_out/r02p_mlp.ncu-rep gpurun_out/r02p_bs23.ncu-rep
ls -la gpurun_out
fi
