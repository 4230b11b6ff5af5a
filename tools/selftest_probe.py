import ctypes, sys
sys.path.insert(0,'/root/repo')
from opencloth_b200 import _abi
bad = ctypes.c_ulonglong(0)
_abi.check(_abi.load().oc_selftest_math(1 << 26, 7, ctypes.byref(bad)))
v = bad.value
print("raw", hex(v))
names = ["scalar(0..19)", "p_add", "p_mul", "p_fma", "sqrt2", "rcp2", "div2", "spring2"]
print("scalar-part", v & ((1<<20)-1))
for i, n in enumerate(names[1:]):
    print(n, (v >> (20 + 4*i)) & 0xf if i < 6 else v >> (20+4*i))
