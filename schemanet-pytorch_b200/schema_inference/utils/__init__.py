from .ingredient_model_wrapper import IngredientModelWrapper

__all__ = ["IngredientModelWrapper"]
