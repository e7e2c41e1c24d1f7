"""Does this box expose NVSwitch multicast (cuMulticastCreate / multimem) to user processes?  One line per GPU."""
import torch
from cuda.bindings import driver as cu

torch.cuda.init()
n = torch.cuda.device_count()
for i in range(n):
    err, dev = cu.cuDeviceGet(i)
    vals = {}
    for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED",
                 "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED"):
        a = getattr(cu.CUdevice_attribute, name, None)
        if a is None:
            vals[name] = "n/a"
            continue
        err, v = cu.cuDeviceGetAttribute(a, dev)
        vals[name.replace("CU_DEVICE_ATTRIBUTE_", "").lower()] = (int(v) if err == cu.CUresult.CUDA_SUCCESS else str(err))
    print(f"gpu {i}: {vals}")
if n >= 2:
    prop = cu.CUmulticastObjectProp()
    prop.numDevices = n
    prop.size = 1 << 29
    prop.handleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    err, gran = cu.cuMulticastGetGranularity(prop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED)
    print("multicast granularity:", err, gran)
    err, h = cu.cuMulticastCreate(prop)
    print("cuMulticastCreate:", err)
