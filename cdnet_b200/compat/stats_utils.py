"""stats_utils.py of the reference: the instance metrics (:7-98, 182-275, 323-334)."""
from ..metrics import (get_dice_1, get_dice_2, get_fast_aji, get_fast_aji_plus, get_fast_dice_2, get_fast_pq,  # noqa: F401
                       remap_label)
