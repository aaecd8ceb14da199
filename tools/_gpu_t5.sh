mkdir -p gpurun_out
python tools/cnx_error_probe.py > gpurun_out/cnx_probe.log 2>&1; cat gpurun_out/cnx_probe.log
for pf in 0 1 2 3; do echo "EFFOCR_PLN_PREFETCH=$pf"; EFFOCR_PLN_PREFETCH=$pf python tools/ab_kernels.py proj_ln; done
python tools/ab_kernels.py
timeout 600 python -m pytest tests/test_gpu_transcription.py tests/test_gpu_recognizer.py tests/test_gpu_blocks.py -q -rf 2>&1 | tail -5
