"""GRU regressor of BASELINE config 5 (reference examples/pydynet/ts_prediction.py:53-69): batch-first single-layer GRU,
Linear head on the last hidden state, MSE loss."""
import numpy as np

import pydynet_b200.nn as nn
import pydynet_b200.nn.functional as F


class GRURegressor(nn.Module):

    def __init__(self, input_size, hidden_size, dtype=np.float32):
        super().__init__()
        self.rnn = nn.GRU(input_size=input_size, hidden_size=hidden_size, num_layers=1, batch_first=True, dtype=dtype)
        self.out = nn.Linear(hidden_size, 1, dtype=dtype)

    def forward(self, x, h_state=None):
        _, h_state = self.rnn(x, h_state)
        return self.out(h_state[:, self.rnn.num_layers - 1, :])


def train_step(net, optimizer, X, Y):
    loss = F.mse_loss(net(X, None), Y)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss
