# A/B timing of the decoder's variants (knobs read when a context is created, bp_decode.cu) + optional parity run; output in
# gpurun_out/.  KML_DEC_DIAG selects timing ablations whose results are wrong (no barriers / no MUFU).
mkdir -p gpurun_out
if [ "$1" = "parity" ]; then python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/ab_parity.log; fi
( for v in "" "KML_DEC_RATIO=2" "KML_DEC_RATIO=3" "KML_DEC_PLANAR=1" "KML_DEC_MINB=2" "KML_DEC_MINB=4" "KML_DEC_DIAG=1" "KML_DEC_DIAG=2"; do
    echo "== $v"; env $v python tools/prof_decode.py 16384 5 -5 | head -1; done
  for v in "" "KML_DEC_T8064=672" "KML_DEC_RATIO=2"; do echo "== PEG8064 $v"; env $v python tools/prof_decode.py 4096 5 5 PEG8064regular0.5.txt 6bits_64QAM_Gray.txt | head -1; done
  for v in "" "KML_DEC_NO_QC=1" "KML_ALG=1" "KML_ALG=1 KML_DEC_NO_QC=1"; do echo "== 5G $v"; env $v python tools/prof_decode.py 8192 5 0 5GLDPCBG2a3_R12_K960.txt 4bit_16QAM_Gray.txt | head -1; done ) 2>&1 | tee gpurun_out/ab_variants.log
