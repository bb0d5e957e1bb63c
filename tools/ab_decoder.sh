set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/ab1_parity.log
cat gpurun_out/ab1_parity.log
( for v in "" "KML_DEC_PLANAR=1" "KML_DEC_RATIO=0" "KML_DEC_RATIO=1" "KML_DEC_PLANAR=1 KML_DEC_RATIO=0" "KML_DEC_MINB=4" "KML_DEC_MINB=2"; do echo "== $v"; env $v python tools/prof_decode.py 16384 5 -5 | head -1; done
echo "== PEG8064 rowmajor"; python tools/prof_decode.py 4096 5 5 PEG8064regular0.5.txt 6bits_64QAM_Gray.txt | head -1
echo "== PEG8064 planar"; KML_DEC_PLANAR=1 python tools/prof_decode.py 4096 5 5 PEG8064regular0.5.txt 6bits_64QAM_Gray.txt | head -1
echo "== 5G"; python tools/prof_decode.py 8192 5 0 5GLDPCBG2a3_R12_K960.txt 4bit_16QAM_Gray.txt | head -1 ) 2>&1 | tee gpurun_out/ab1_variants.log
