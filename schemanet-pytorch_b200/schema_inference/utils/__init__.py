"""Hot-path part of the reference's `schema_inference.utils`: only the ingredient wrapper lives here (checkpoint
loaders, DDP helpers and parameter-group utilities are out of scope)."""
from .ingredient_model_wrapper import IngredientModelWrapper

__all__ = ("IngredientModelWrapper",)
