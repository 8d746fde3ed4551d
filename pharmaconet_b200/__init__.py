"""pharmaconet_b200 - B200-native implementation of PharmacoNet's virtual-screening hot path.

`from pharmaconet_b200 import PharmacophoreModel` mirrors `from pmnet import PharmacophoreModel`
(src/pmnet/__init__.py:1); the CNN side lives in `pharmaconet_b200.module` (`PharmacoNet`, `get_pmnet_dev`)."""

from .pharmacophore_model import PharmacophoreModel

__version__ = "0.1.0"
__all__ = ["PharmacophoreModel"]
