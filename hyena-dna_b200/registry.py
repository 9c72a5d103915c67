"""Name -> class registry for the two layers on the hot path (src/utils/registry.py:40-41)."""
from .hyena import HyenaFilter, HyenaOperator

layer = {
    "hyena": HyenaOperator,
    "hyena-filter": HyenaFilter,
}
