# A/B timing of the decoder's compile-time variants (env knobs in bp_decode.cu) + optional parity run; output in gpurun_out/
mkdir -p gpurun_out
if [ "$1" = "parity" ]; then python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/ab_parity.log; fi
( for v in "" "KML_DEC_RATIO=2" "KML_DEC_T=576" "KML_DEC_T=576 KML_DEC_RATIO=2" "KML_DEC_PLANAR=1"; do echo "== $v"; env $v python tools/prof_decode.py 16384 5 -5 | head -1; done
echo "== PEG8064"; python tools/prof_decode.py 4096 5 5 PEG8064regular0.5.txt 6bits_64QAM_Gray.txt | head -1
echo "== PEG8064 RATIO=2"; KML_DEC_RATIO=2 python tools/prof_decode.py 4096 5 5 PEG8064regular0.5.txt 6bits_64QAM_Gray.txt | head -1
echo "== 5G"; python tools/prof_decode.py 8192 5 0 5GLDPCBG2a3_R12_K960.txt 4bit_16QAM_Gray.txt | head -1 ) 2>&1 | tee gpurun_out/ab_variants.log
