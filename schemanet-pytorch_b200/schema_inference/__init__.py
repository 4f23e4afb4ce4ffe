"""Drop-in for the hot-path part of the reference's `schema_inference` package: `schema_inference.graph`
(SchemaNet, Matcher, GNN, SchemaNetPredictor) and `schema_inference.utils.IngredientModelWrapper`.
Training workers, evaluation loops, losses and data pipelines of the reference are out of scope."""
