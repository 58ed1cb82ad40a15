"""Is there a Vulkan loader + ICD on this machine that could run the reference's own compute pass?

TEST / BENCH INFRASTRUCTURE. The reference renders through Vulkan
(src/rvpt/rvpt.cpp:646-655 binds the shader's descriptors, :1035-1036 dispatches it); with a
software ICD (lavapipe) its shipped compute_pass.comp.spv would run on the host cores and be the
CPU baseline of kind "reference" in the strict sense. This image has neither loader, ICD nor
headers, so the baseline is the same shader binary translated to C++ (oracle/spirv_to_cpp.py);
this probe only records, in every reference-arm bench line, whether that is still the situation
on the machine the bench ran on. It loads the loader if there is one, asks for the instance
version and counts the physical devices — nothing else.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util


def probe() -> dict:
    name = ctypes.util.find_library("vulkan") or "libvulkan.so.1"
    try:
        vk = C.CDLL(name)
    except OSError as e:
        return {"available": False, "why": f"no Vulkan loader ({name}: {e})".replace("\n", " ")[:160]}
    out = {"available": False, "loader": name}
    try:
        version = C.c_uint32(0)
        if hasattr(vk, "vkEnumerateInstanceVersion") and vk.vkEnumerateInstanceVersion(C.byref(version)) == 0:
            v = version.value
            out["instance_version"] = f"{v >> 22 & 0x7f}.{v >> 12 & 0x3ff}.{v & 0xfff}"

        class AppInfo(C.Structure):  # VkApplicationInfo
            _fields_ = [("sType", C.c_int32), ("pNext", C.c_void_p), ("pApplicationName", C.c_char_p),
                        ("applicationVersion", C.c_uint32), ("pEngineName", C.c_char_p),
                        ("engineVersion", C.c_uint32), ("apiVersion", C.c_uint32)]

        class InstanceCreateInfo(C.Structure):  # VkInstanceCreateInfo
            _fields_ = [("sType", C.c_int32), ("pNext", C.c_void_p), ("flags", C.c_uint32),
                        ("pApplicationInfo", C.POINTER(AppInfo)), ("enabledLayerCount", C.c_uint32),
                        ("ppEnabledLayerNames", C.c_void_p), ("enabledExtensionCount", C.c_uint32),
                        ("ppEnabledExtensionNames", C.c_void_p)]

        app = AppInfo(0, None, b"rvpt_b200 probe", 0, None, 0, 1 << 22)   # VK_STRUCTURE_TYPE_APPLICATION_INFO, 1.0
        info = InstanceCreateInfo(1, None, 0, C.pointer(app), 0, None, 0, None)  # ..._INSTANCE_CREATE_INFO
        inst = C.c_void_p()
        rc = vk.vkCreateInstance(C.byref(info), None, C.byref(inst))
        if rc != 0:
            out["why"] = f"vkCreateInstance returned {rc} (no usable ICD)"
            return out
        n = C.c_uint32(0)
        vk.vkEnumeratePhysicalDevices(inst, C.byref(n), None)
        out["physical_devices"] = n.value
        vk.vkDestroyInstance(inst, None)
        out["available"] = n.value > 0
        if not out["available"]:
            out["why"] = "loader present, no physical device (no ICD)"
    except Exception as e:  # a broken loader must not break a bench line
        out["why"] = f"probe failed: {e}"[:160]
    return out


if __name__ == "__main__":
    print(probe())
