"""mettagrid_b200 -- B200-native batched MettaGrid step."""

from .compiler import CompileError, Program, compile_config  # noqa: F401

__version__ = "0.1.0"
