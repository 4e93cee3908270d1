"""`_elf`: the names the reference's `elf` Python package imports from its pybind module
(`from _elf import *`, src_py/elf/__init__.py:8; exported by src_cpp/elf/Pybind.cc:28-88,91-117),
served by elf_b200.compat.  Submodules `_options` and `_logging` as in Pybind.cc:91-117."""
from elf_b200.compat import (AnyP, Context, ContextOptions, ReplyStatus, SearchAlgoOptions, SharedMem,  # noqa: F401
                             SharedMemOptions, Size, TSOptions)

from . import _logging, _options  # noqa: F401


class FuncMapBase:
    """elf::FuncMapBase (extractor.h:262-300) is only ever reached through AnyP.field() from Python;
    exported for `from _elf import *` parity."""


__all__ = ["Context", "SharedMem", "SharedMemOptions", "AnyP", "FuncMapBase", "Size", "ReplyStatus", "TSOptions",
           "SearchAlgoOptions", "ContextOptions", "_logging", "_options"]
