from .sam import Sam  # noqa: F401
from .image_encoder import ImageEncoderViT  # noqa: F401
from .pos_encoder import PromptEncoder  # noqa: F401
from .transformer import TwoWayTransformer  # noqa: F401
from .box_decoder import MaskDecoder  # noqa: F401
