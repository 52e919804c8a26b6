"""`mock` (PyPI backport) stand-in for the reference's tests: re-export the stdlib implementation."""
from unittest.mock import *  # noqa: F401,F403
from unittest.mock import MagicMock, Mock, call, patch  # noqa: F401
