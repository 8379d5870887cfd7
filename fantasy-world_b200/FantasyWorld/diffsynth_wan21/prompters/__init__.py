from .wan_prompter import WanPrompter
