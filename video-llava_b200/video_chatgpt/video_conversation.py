"""Prompt templates (reference: video_chatgpt/video_conversation.py:14-179), restated for the one
template family the inference path uses: system + sep + 'ROLE: message' turns, two separators."""
import dataclasses
from enum import Enum, auto
from typing import List


class SeparatorStyle(Enum):
    SINGLE = auto()
    TWO = auto()


@dataclasses.dataclass
class Conversation:
    system: str
    roles: List[str]
    messages: List[List[str]]
    offset: int
    sep_style: SeparatorStyle = SeparatorStyle.SINGLE
    sep: str = "###"
    sep2: str = None
    version: str = "Unknown"

    def get_prompt(self):
        if self.sep_style == SeparatorStyle.SINGLE:
            out = self.system + self.sep
            for role, msg in self.messages:
                if msg:
                    out += role + ": " + (msg[0] if isinstance(msg, tuple) else msg) + self.sep
                else:
                    out += role + ":"
            return out
        if self.sep_style == SeparatorStyle.TWO:
            seps = [self.sep, self.sep2]
            out = self.system + seps[0]
            for i, (role, msg) in enumerate(self.messages):
                if msg:
                    out += role + ": " + (msg[0] if isinstance(msg, tuple) else msg) + seps[i % 2]
                else:
                    out += role + ":"
            return out
        raise ValueError(f"Invalid style: {self.sep_style}")

    def append_message(self, role, message):
        self.messages.append([role, message])

    def copy(self):
        return Conversation(system=self.system, roles=self.roles, messages=[[r, m] for r, m in self.messages],
                            offset=self.offset, sep_style=self.sep_style, sep=self.sep, sep2=self.sep2,
                            version=self.version)


def _assistant_system(name: str) -> str:
    # system prompt text of the reference templates (video_conversation.py:145-148,159-162); the
    # missing space before "Follow" is the reference's and is part of the prompt bytes
    return (f"You are {name}, a large vision-language assistant. "
            "You are able to understand the video content that the user provides, and assist the user with a "
            "variety of tasks using natural language."
            "Follow the instructions carefully and explain your answers in detail based on the provided video.")


_VICUNA_SYSTEM = ("A chat between a curious user and an artificial intelligence assistant. "
                  "The assistant gives helpful, detailed, and polite answers to the user's questions.")


def _two(system):
    return Conversation(system=system, roles=("USER", "ASSISTANT"), version="v1", messages=[], offset=0,
                        sep_style=SeparatorStyle.TWO, sep=" ", sep2="</s>")


conv_vicuna_v1_1 = _two(_VICUNA_SYSTEM)
conv_video_chatgpt_v1 = _two(_assistant_system("Video-ChatGPT"))
conv_pg_video_llava = _two(_assistant_system("PG-Video-LLaVA"))

# The reference's "default" is a SINGLE-style few-shot template that no inference caller selects;
# the inference entry points always pass conv_mode explicitly ("pg-video-llava" / "video-chatgpt_v1").
default_conversation = conv_pg_video_llava
conv_templates = {"video-chatgpt_v1": conv_video_chatgpt_v1, "vicuna_v1_1": conv_vicuna_v1_1,
                  "pg-video-llava": conv_pg_video_llava}
