"""`import grassmanntn as gtn` -- drop-in alias of the B200 package (grassmanntn_b200) under the reference's own name,
so that scripts written against ayosprakob/grassmanntn (`import grassmanntn as gtn`, `from grassmanntn import param`,
`from grassmanntn import gauge2d_block as gauge`) run unchanged.  Put the repository root on sys.path (instead of the
reference checkout) to switch."""
import sys

import grassmanntn_b200 as _impl

for _sub in ("param", "gauge2d", "gauge2d_block", "checkpoint", "parallel", "sharded"):
    sys.modules[__name__ + "." + _sub] = getattr(_impl, _sub, None) or __import__("grassmanntn_b200." + _sub, fromlist=[_sub])
sys.modules[__name__] = _impl
