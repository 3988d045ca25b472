"""Drop-in for the reference's ``active_selection/my_bvsb_predclsbal_pwr_banignore.py`` (same module and class name)."""
from . import base


class RegionSelector(base.RegionSelector):
    method_name = "my_bvsb_predclsbal_pwr_banignore"
