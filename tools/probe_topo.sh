nvidia-smi topo -m 2>&1 | head -30
echo ---; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" 
echo ---; ls /sys/devices/system/node/ 2>&1 | head; cat /sys/devices/system/node/node*/cpulist 2>/dev/null
echo ---; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class; then echo $d $(cat $d/numa_node 2>/dev/null) $(cat $d/local_cpulist 2>/dev/null); fi; done
echo ---; nproc; python - <<'PY'
import os; print(sorted(os.sched_getaffinity(0)))
import torch
for i in range(torch.cuda.device_count()):
    p=torch.cuda.get_device_properties(i); print(i, p.name, getattr(p,'pci_bus_id',None), getattr(p,'pci_domain_id',None), getattr(p,'pci_device_id',None))
PY
nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv
