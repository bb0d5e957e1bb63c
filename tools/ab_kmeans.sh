for v in base minb3 minb5; do
  if [ $v != base ]; then cp tools/_bin/libkm_$v.so kmldpc_b200/lib/libkmldpc_b200.so; fi
  echo "== $v"; timeout 60 python tools/prof_kmeans.py 15; timeout 60 python tools/prof_kmeans.py -5
done
