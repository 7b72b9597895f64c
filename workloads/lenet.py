"""LeNet-style ConvNet of BASELINE config 2 (reference examples/pydynet/mnist.py:82-98): 1x28x28 -> conv3x3(20) -> relu ->
pool2 -> conv3x3(50) -> relu -> pool2 -> fc 2450-500 -> relu -> fc 500-10."""
import numpy as np

import pydynet_b200 as pdn
import pydynet_b200.nn as nn
import pydynet_b200.nn.functional as F


class ConvNet(nn.Module):

    def __init__(self, dtype=np.float32):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 20, 3, 1, 1, dtype=dtype)
        self.conv2 = nn.Conv2d(20, 50, 3, 1, 1, dtype=dtype)
        self.fc1 = nn.Linear(7 * 7 * 50, 500, dtype=dtype)
        self.fc2 = nn.Linear(500, 10, dtype=dtype)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.conv1(x)), 2, 2)
        x = F.max_pool2d(F.relu(self.conv2(x)), 2, 2)
        x = F.relu(self.fc1(x.reshape(-1, 7 * 7 * 50)))
        return self.fc2(x)


def train_step(net, optimizer, X, y):
    """One step of reference examples/pydynet/mnist.py:161-166: CE loss -> zero_grad -> backward -> step."""
    loss = F.cross_entropy_loss(net(X), y)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss
