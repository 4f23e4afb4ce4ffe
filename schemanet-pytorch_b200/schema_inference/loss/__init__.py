"""Drop-in for `schema_inference.loss` (schema_inference/loss/__init__.py): the loss registry of the training workflow."""
from typing import Any, Dict

from .schema_inference_loss import CELoss, Loss, SchemaInferenceLoss

__REGISTERED_LOSS__ = {"ce_loss": CELoss, "schema_inference_loss": SchemaInferenceLoss}


def get_loss_fn(loss_cfg: Dict[str, Any], **kwargs) -> Loss:
    return __REGISTERED_LOSS__[loss_cfg["name"]](**loss_cfg.get("loss_cfg", dict()), **kwargs)
