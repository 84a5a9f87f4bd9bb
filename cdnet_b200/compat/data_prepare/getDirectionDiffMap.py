"""data_prepare/getDirectionDiffMap.py of the reference: `circshift` (:14-42), `generate_dd_map` (:44-108)."""
from ...api import circshift, generate_dd_map  # noqa: F401
