"""feddrift_b200 — a Blackwell-native federated-learning simulation engine with the
capabilities of microsoft/FedDrift (FedML programming model + concept-drift FL
algorithms), designed for 8×B200: device-resident parameter arena, fused
sm_100a kernels for local step / aggregation / evaluation, CUDA-graph round
loop, NVLink peer-memory aggregation.  See DESIGN.md."""
__version__ = "0.1.0"
