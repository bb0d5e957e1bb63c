cp tools/_bin/lib_tuning.so kmldpc_b200/lib/libkmldpc_b200.so
for c in "" "KML_RX_CHUNK=4096" "KML_RX_CHUNK=4096 KML_RX_SLOW=0" "KML_RX_CHUNK=8192 KML_RX_SLOW=0" "KML_RX_CHUNK=16384 KML_RX_SLOW=0" "KML_RX_CHUNK=2048 KML_RX_SLOW=0" "KML_RX_CHUNK=1024 KML_RX_SLOW=0"; do
  env $c timeout 100 python tools/ab_e2e.py 2>/dev/null | tail -1
done
