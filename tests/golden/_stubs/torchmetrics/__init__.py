"""Import stub (golden generation only): the reference's TrainingMetrics derives from torchmetrics.Metric."""
