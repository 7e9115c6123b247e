"""Drop-in counterparts of the reference's ``architectures`` package (generator_with_attention.py,
discriminator_with_attention.py): same class names, constructor and build-method signatures, same
tensor attributes -- backed by the sm_100a kernels behind the C ABI instead of a TensorFlow graph."""
from .discriminator_with_attention import Discriminator  # noqa: F401
from .generator_with_attention import Generator  # noqa: F401
