LANGUAGES = {"en": "english"}
