"""my_transforms_direction.py of the reference: `get_centerpoint2` (:651-685), `LabelEncoding` (:687-885)."""
from ..api import LabelEncoding, get_centerpoint2  # noqa: F401
