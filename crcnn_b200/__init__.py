"""crcnn_b200: B200-native engine for CrCNN's encrypted-inference forward pass.

The product is the C-ABI shared library (include/crcnn_b200.h, built from crcnn_b200/csrc/) and the
C++17 layer classes in crcnn_b200/cpp/.  The Python modules here are harness plumbing only."""
