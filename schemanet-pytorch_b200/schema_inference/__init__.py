"""Drop-in for the hot-path part of the reference's `schema_inference` package: `schema_inference.graph`
(SchemaNet, Matcher, GNN, SchemaNetPredictor), `schema_inference.utils.IngredientModelWrapper` and, for the training
step of SURVEY.md section 8 f3, `schema_inference.loss`.  Training workers, evaluation loops and data pipelines of the
reference are out of scope."""
