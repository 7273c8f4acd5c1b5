"""Training-side pieces of the BACS path: loss modules, the IoU metric, the replay buffer."""
