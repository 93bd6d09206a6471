mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash profiles/run_ab.sh "-DNFE_REC_STAGE=0" "" > gpurun_out/ab_recstage.txt 2>&1
cat gpurun_out/ab_recstage.txt
NFE_NVCC_FLAGS="-DNFE_PIPE_PROFILE" python -m nerffaceediting_b200.build --force > /dev/null 2>&1
python profiles/pipe_role_profile.py > gpurun_out/role_profile.txt 2>&1
cat gpurun_out/role_profile.txt
