"""Model definitions of the BASELINE configs, written against the pydynet_b200 API (same attribute / parameter names as
the reference's own model files so weights map one-to-one):
  lenet.ConvNet          reference examples/pydynet/mnist.py:82-98        (config 2)
  llama.Llama            reference llm/llama/model.py:10-269              (config 3)
  encoder.Transformer    reference examples/pydynet/transformer.py:53-192 (config 4)
  gru.GRURegressor       reference examples/pydynet/ts_prediction.py:53-69 (config 5)
"""
