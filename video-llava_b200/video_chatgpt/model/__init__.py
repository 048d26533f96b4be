from .video_chatgpt import (CLIPVisionTower, VideoChatGPTConfig, VideoChatGPTLlamaForCausalLM,  # noqa: F401
                            VideoChatGPTLlamaModel, VisionConfig)
