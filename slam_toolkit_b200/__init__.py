"""Import shim: the product package lives in ``slam-toolkit_b200/`` (a name Python
cannot import directly); this package forwards its search path there."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "slam-toolkit_b200"))
