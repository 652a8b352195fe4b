"""Reference-side bindings of libmsplat_b200.so.

``integration._C`` is a drop-in for the reference's pybind11 module ``msplat._C``
(/root/reference/msplat/src/ext.cpp:14-25): the same 12 function names, argument orders and return
tuples, implemented over the C ABI of include/msplat_b200.h.  ``integration.load`` runs the reference's
UNMODIFIED Python wrappers (msplat/*.py) on top of it.
"""
