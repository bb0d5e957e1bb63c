# host topology as seen from the GPU box's container (for the e2e host->device limit at N = 8)
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"
echo "cpus allowed: $(grep Cpus_allowed_list /proc/self/status)"; grep Mems_allowed_list /proc/self/status
cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
nvidia-smi topo -m 2>&1 | head -20
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class)" = "0x030200" ]; then echo "$d numa $(cat $d/numa_node) cpus $(cat $d/local_cpulist)"; fi; done
ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist) $(grep MemFree $n/meminfo)"; done
python - <<'PY'
import ctypes, os
print("affinity", sorted(os.sched_getaffinity(0)))
PY
